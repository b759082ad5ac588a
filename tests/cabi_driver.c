/* cabi_driver.c -- the drop-in boundary exercised from plain C, with no Python in the process: what a Rust / C / C++ host
 * does with libyasph_gpu.so.  Builds the application's dam-break scene (main.rs:177-196) with the ABI's scene builders and
 *   cabi_driver scene          prints the particle counts and a hash of the scene (no device needed)
 *   cabi_driver run <steps>    steps it on the GPU through yasph_step_host and prints, per step, dt and iteration counts, and at
 *                              the end a hash of positions / velocities / densities
 * tests/test_cabi_driver.py compares the output with the same run through the ctypes mirror.
 * Compile: gcc -std=c99 -Iinclude tests/cabi_driver.c -Lyasph2d_b200 -lyasph_gpu -Wl,-rpath,$PWD/yasph2d_b200 -o cabi_driver */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "yasph_gpu.h"

static uint64_t fnv1a(const void* data, size_t bytes, uint64_t h) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < bytes; ++i) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

typedef struct {
    float* xy;
    uint32_t n, cap;
} Cloud;

static void reserve(Cloud* c, uint32_t extra) {
    if (c->n + extra > c->cap) {
        c->cap = (c->n + extra) * 2 + 1024;
        c->xy = (float*)realloc(c->xy, (size_t)c->cap * 2 * sizeof(float));
    }
}
static int thick_line(Cloud* c, float pd, float sx, float sy, float ex, float ey, uint32_t thickness) {
    uint32_t cnt = 0;
    if (yasph_scene_boundary_thick_line(pd, sx, sy, ex, ey, thickness, NULL, 0, &cnt)) return 1;
    reserve(c, cnt);
    if (yasph_scene_boundary_thick_line(pd, sx, sy, ex, ey, thickness, c->xy + 2 * (size_t)c->n, cnt, &cnt)) return 1;
    c->n += cnt;
    return 0;
}

int main(int argc, char** argv) {
    const float particle_density = 10000.0f;
    Cloud fluid = {0, 0, 0}, wall = {0, 0, 0};
    uint32_t cnt = 0;
    /* main.rs:180: add_fluid_rect(Rect(0.1, 0.7, 0.5, 1.0), 0.05), seeded with the particle count so far (0) */
    if (yasph_scene_fluid_rect(particle_density, 0.1f, 0.7f, 0.5f, 1.0f, 0.05f, 0, NULL, 0, &cnt)) return 2;
    reserve(&fluid, cnt);
    if (yasph_scene_fluid_rect(particle_density, 0.1f, 0.7f, 0.5f, 1.0f, 0.05f, 0, fluid.xy, cnt, &cnt)) return 2;
    fluid.n = cnt;
    /* main.rs:182-195 */
    if (thick_line(&wall, particle_density, 0.0f, 2.5f, 2.0f, 2.5f, 4) || thick_line(&wall, particle_density, 0.0f, 0.0f, 2.0f, 0.0f, 4) ||
        thick_line(&wall, particle_density, 0.0f, 0.0f, 0.0f, 2.5f, 4) || thick_line(&wall, particle_density, 2.0f, 0.0f, 2.0f, 2.5f, 4) ||
        thick_line(&wall, particle_density, 0.0f, 0.6f, 1.75f, 0.5f, 2) || thick_line(&wall, particle_density, 0.0f, 2.5f, 2.0f, 2.5f, 2) ||
        thick_line(&wall, particle_density, -2.0f, -0.5f, 4.0f, -0.5f, 4))
        return 2;
    uint64_t h = fnv1a(fluid.xy, (size_t)fluid.n * 8, 14695981039346656037ull);
    h = fnv1a(wall.xy, (size_t)wall.n * 8, h);
    printf("scene fluid=%u boundary=%u hash=%016llx\n", fluid.n, wall.n, (unsigned long long)h);
    if (argc < 2 || strcmp(argv[1], "run") != 0) return 0;

    const int steps = argc > 2 ? atoi(argv[2]) : 10;
    yasph_config cfg;
    if (yasph_config_default(&cfg, 2.0f, particle_density, 100.0f, YASPH_SOLVER_DFSPH)) return 3;
    cfg.max_particles = fluid.n;
    cfg.max_boundary = wall.n;
    yasph_ctx* ctx = NULL;
    if (yasph_create(&cfg, &ctx) != YASPH_OK) {
        fprintf(stderr, "yasph_create: %s\n", yasph_last_error(NULL));
        return 4;
    }
    float* vel = (float*)calloc((size_t)fluid.n * 2, sizeof(float));
    float* dens = (float*)calloc(fluid.n, sizeof(float));
    int rc = yasph_set_boundary(ctx, wall.xy, wall.n);
    for (int s = 0; s < steps && rc == YASPH_OK; ++s) {
        yasph_step_report rep;
        rc = yasph_step_host(ctx, fluid.xy, vel, dens, fluid.n, &rep); /* Solver::simulation_step on the host's own arrays */
        if (rc == YASPH_OK)
            printf("step %d dt_ns=%llu iters=%u/%u\n", s, (unsigned long long)rep.dt_ns, rep.iters_density, rep.iters_divergence);
    }
    if (rc != YASPH_OK) {
        fprintf(stderr, "error %d: %s\n", rc, yasph_last_error(ctx));
        return 5;
    }
    h = fnv1a(fluid.xy, (size_t)fluid.n * 8, 14695981039346656037ull);
    h = fnv1a(vel, (size_t)fluid.n * 8, h);
    h = fnv1a(dens, (size_t)fluid.n * 4, h);
    printf("state hash=%016llx\n", (unsigned long long)h);
    yasph_destroy(ctx);
    free(vel);
    free(dens);
    free(fluid.xy);
    free(wall.xy);
    return 0;
}
