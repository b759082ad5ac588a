// scene.inl -- host-side scene builders behind the C ABI (no device work).  Mirrors FluidParticleWorld::add_fluid_rect /
// add_boundary_line / add_boundary_thick_line (src/sph/fluidparticleworld.rs:140-195) so that benchmarks and the host
// mirrors can build the reference's scenes without touching oracle/.
#include <algorithm>
#include <cmath>

namespace {
// rand 0.8 SmallRng on 64-bit targets: xoshiro256++, state filled by SplitMix64 (seed_from_u64)
struct SceneRng {
    uint64_t s[4];
    explicit SceneRng(uint64_t seed) {
        for (int i = 0; i < 4; ++i) {
            seed += 0x9e3779b97f4a7c15ull;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
            s[i] = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    float gen_f32() { return (float)((uint32_t)(next() >> 32) >> 8) * (1.0f / 16777216.0f); }
};
struct P2 {
    float x, y;
};
inline P2 add(P2 a, P2 b) { return P2{a.x + b.x, a.y + b.y}; }
inline P2 sub(P2 a, P2 b) { return P2{a.x - b.x, a.y - b.y}; }
inline P2 mul(P2 a, float s) { return P2{a.x * s, a.y * s}; }
inline P2 dvd(P2 a, float s) { return P2{a.x / s, a.y / s}; }

uint32_t boundary_line(float nppm, P2 start, P2 end, float* out, uint32_t cap, uint32_t at) {
    const P2 d = sub(end, start);
    const float distance = sqrtf(d.x * d.x + d.y * d.y);
    const size_t n = std::max<size_t>(1, (size_t)ceilf(distance * nppm));
    const P2 step = dvd(dvd(d, distance), nppm);
    P2 pos = start;
    for (size_t i = 0; i < n; ++i) {
        if (out && at + i < cap) {
            out[2 * (at + i)] = pos.x;
            out[2 * (at + i) + 1] = pos.y;
        }
        pos = add(pos, step);
    }
    return (uint32_t)n;
}
}  // namespace

extern "C" int32_t yasph_scene_fluid_rect(float particle_density, float x, float y, float w, float h, float jitter, uint64_t seed,
                                          float* out_xy, uint32_t capacity, uint32_t* count) {
    if (!count || !(particle_density > 0.f)) return YASPH_ERR_INVALID_ARGUMENT;
    const float nppm = sqrtf(particle_density) * 0.9f;
    const size_t nx = std::max<size_t>(1, (size_t)(w * nppm));
    const size_t ny = std::max<size_t>(1, (size_t)(h * nppm));
    *count = (uint32_t)(nx * ny);
    if (!out_xy) return YASPH_OK;
    if (capacity < nx * ny) return YASPH_ERR_CAPACITY;
    SceneRng rng(seed);
    const float step = fminf(w / (float)nx, h / (float)ny);
    const float jf = step * jitter;
    size_t k = 0;
    for (size_t iy = 0; iy < ny; ++iy)
        for (size_t ix = 0; ix < nx; ++ix, ++k) {
            const float jx = rng.gen_f32(), jy = rng.gen_f32();
            const P2 jit = mul(add(mul(P2{jx, jy}, 0.5f), P2{0.5f, 0.5f}), jf);
            const P2 p = add(add(P2{x, y}, jit), P2{step * (float)ix, step * (float)iy});
            out_xy[2 * k] = p.x;
            out_xy[2 * k + 1] = p.y;
        }
    return YASPH_OK;
}
extern "C" int32_t yasph_scene_boundary_line(float particle_density, float sx, float sy, float ex, float ey, float* out_xy,
                                             uint32_t capacity, uint32_t* count) {
    if (!count || !(particle_density > 0.f)) return YASPH_ERR_INVALID_ARGUMENT;
    *count = boundary_line(sqrtf(particle_density), P2{sx, sy}, P2{ex, ey}, nullptr, 0, 0);
    if (!out_xy) return YASPH_OK;
    if (capacity < *count) return YASPH_ERR_CAPACITY;
    boundary_line(sqrtf(particle_density), P2{sx, sy}, P2{ex, ey}, out_xy, capacity, 0);
    return YASPH_OK;
}
extern "C" int32_t yasph_scene_boundary_thick_line(float particle_density, float sx, float sy, float ex, float ey, uint32_t thickness,
                                                   float* out_xy, uint32_t capacity, uint32_t* count) {
    if (!count || !(particle_density > 0.f) || thickness == 0) return YASPH_ERR_INVALID_ARGUMENT;
    const float nppm = sqrtf(particle_density);
    const P2 start{sx, sy}, end{ex, ey};
    const P2 d = sub(end, start);
    const P2 dir = mul(d, 1.0f / sqrtf(d.x * d.x + d.y * d.y));  // cgmath normalize
    const P2 perp{-dir.y, dir.x};
    const float tw = (float)thickness / nppm;
    const P2 elong = mul(dir, tw);
    P2 offset = mul(P2{-perp.x, -perp.y}, tw);
    const P2 step = dvd(mul(perp, tw), (float)thickness);
    uint32_t total = 0;
    for (int pass = 0; pass < 2; ++pass) {
        P2 off = offset;
        uint32_t at = 0;
        for (uint32_t i = 0; i < thickness; ++i) {
            at += boundary_line(nppm, add(start, off), add(add(end, off), elong), pass ? out_xy : nullptr, capacity, at);
            off = add(off, step);
        }
        total = at;
        if (!out_xy) break;
        if (pass == 0 && capacity < total) {
            *count = total;
            return YASPH_ERR_CAPACITY;
        }
    }
    *count = total;
    return YASPH_OK;
}
extern "C" uint64_t yasph_duration_from_secs_f32(float secs) { return yasph::duration_from_secs_f32(secs); }
extern "C" float yasph_duration_as_secs_f32(uint64_t ns) { return yasph::duration_as_secs_f32(ns); }
