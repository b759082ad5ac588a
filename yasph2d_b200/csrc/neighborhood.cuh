// neighborhood.cuh -- Morton-sorted compact cell grid, 8x8-cell tiles and compact per-particle neighbour lists.
//
// Replaces CompactMortonCellGrid::update (src/sph/neighborhood_search.rs:90-166) and NeighborLists::try_update
// (:312-397).  Semantics kept from the reference: candidates of a particle are all particles in the 3x3 cell box
// of its cell, visited in ascending sorted index (== ascending Morton code of the cell, neighborhood_search.rs:191-259
// produces exactly that order as <=5 index runs); a candidate is a neighbour iff d2 <= r2 && d2 > 1e-10 with
// d2 = fl(fl(dx*dx) + fl(dy*dy)) (:356-357); dynamic neighbours first, then static, at most 64 in total (:322).
//
// B200 design: particles are stored Morton-sorted, so an aligned 8x8 block of cells (a "tile") is one contiguous
// index range.  Every neighbour-dependent pass runs one CTA per tile, stages the tile's particles plus its 1-cell
// apron (<= 100 cells, looked up once per step into a per-tile table) in shared memory in ascending-key order, and
// then addresses neighbours by 16-bit shared-memory slot.  Ascending slot == ascending global index, so list order
// (and with it every floating-point sum order) is the reference's.  Lists are stored as u16 slots, 4 per 8-byte
// word, interleaved across the tile's particles so a warp reads them coalesced: 2 B per neighbour instead of the
// reference's 4 B + 8 B range record (neighborhood_search.rs:268-273,299).
#pragma once
#include "common.cuh"

namespace yasph {

constexpr int NB_THREADS = 256;

struct TileHeader {       // 32 bytes
    uint32_t pstart;      // first particle of the tile (sorted index)
    uint32_t pcount;      // particles in the tile
    uint32_t dyn_total;   // staged dynamic candidates (tile + apron)
    uint32_t stat_total;  // staged boundary candidates
    uint32_t own_lo;      // slot of the tile's first own particle in the staged dynamic array
    uint32_t pad0, pad1, pad2;
};
// Per tile and region cell in RANK order (ascending Morton key): .x = first global index, .y = slot_start | count << 16
typedef uint2 TileCell;

struct TileTables {
    const TileHeader* hdr;
    const TileCell* dyn;          // [tile][100]
    const TileCell* stat;         // [tile][100]
    const uint8_t* rank_of_cell;  // [tile][128]: region cell (ly*10+lx) -> rank
};

// ---- keys ---------------------------------------------------------------------------------------------------------
// neighborhood_search.rs:111-114 (sequential in the reference)
__global__ void k_keygen(const float2* __restrict__ pos, uint32_t n, GridParams g, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        keys[i] = position_to_cidx(g, pos[i]);
        idx[i] = i;
    }
}
// dfsph.rs:502-509 (advect) fused with the key generation of the following re-sort
__global__ void k_advect_keygen(float2* __restrict__ pos, const float2* __restrict__ vstar, uint32_t n, const Control* __restrict__ ctl,
                                GridParams g, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float dt = ctl->dt;
        float2 p = pos[i] + vstar[i] * dt;
        pos[i] = p;
        keys[i] = position_to_cidx(g, p);
        idx[i] = i;
    }
}
// wscsph.rs:141-150 (leap frog 1) fused with key generation
__global__ void k_kickdrift_keygen(float2* __restrict__ pos, float2* __restrict__ vel, const float2* __restrict__ acc, uint32_t n,
                                   const Control* __restrict__ ctl, GridParams g, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float dt = ctl->dt_prev;
        float2 v = vel[i] + 0.5f * dt * acc[i];
        float2 p = pos[i] + v * dt;
        vel[i] = v;
        pos[i] = p;
        keys[i] = position_to_cidx(g, p);
        idx[i] = i;
    }
}

// apply_sorting (neighborhood_search.rs:71-78): out[k] = in[perm[k]] for up to three float2 and two float arrays
struct GatherArgs {
    const float2* in2[3];
    float2* out2[3];
    const float* in1[2];
    float* out1[2];
    int n2, n1;
};
__global__ void k_gather(const uint32_t* __restrict__ perm, uint32_t n, GatherArgs a) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        uint32_t s = perm[k];
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q < a.n2) a.out2[q][k] = a.in2[q][s];
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (q < a.n1) a.out1[q][k] = a.in1[q][s];
    }
}

// ---- cells and tiles: head flags -> scan -> compaction (functors for scan.cuh) ----------------------------------------
// value = cell head (low 32 bits) | tile head (high 32 bits)
struct HeadFlagsIn {
    const uint32_t* keys;
    __device__ __forceinline__ unsigned long long operator()(uint32_t i) const {
        uint32_t k = keys[i];
        uint32_t p = i ? keys[i - 1] : ~k;
        unsigned long long c = (i == 0 || k != p) ? 1ull : 0ull;
        unsigned long long t = (i == 0 || (k >> YASPH_TILE_SHIFT) != (p >> YASPH_TILE_SHIFT)) ? 1ull : 0ull;
        return c | (t << 32);
    }
};
struct HeadCompactOut {
    const uint32_t* keys;
    uint32_t* cell_key;
    uint32_t* cell_start;
    uint32_t* tile_key;
    uint32_t* tile_pstart;
    uint32_t max_tiles;
    __device__ __forceinline__ void operator()(uint32_t i, unsigned long long ex, unsigned long long v) const {
        if (v & 0xFFFFFFFFull) {
            uint32_t c = (uint32_t)(ex & 0xFFFFFFFFull);
            cell_key[c] = keys[i];
            cell_start[c] = i;
        }
        if (v >> 32) {
            uint32_t t = (uint32_t)(ex >> 32);
            if (t < max_tiles) {
                tile_key[t] = keys[i] >> YASPH_TILE_SHIFT;
                tile_pstart[t] = i;
            }
        }
    }
};
// sentinel cell {first_particle = N, cidx = u32::MAX} (neighborhood_search.rs:161-164) and the counts
__global__ void k_finish_cells(const unsigned long long* __restrict__ total, uint32_t n, uint32_t* cell_key, uint32_t* cell_start,
                               uint32_t* tile_pstart, uint32_t max_tiles, Control* ctl, int is_static) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long t = n ? *total : 0ull;
        uint32_t c = (uint32_t)(t & 0xFFFFFFFFull), nt = (uint32_t)(t >> 32);
        cell_key[c] = 0xFFFFFFFFu;
        cell_start[c] = n;
        if (is_static) {
            ctl->num_cells_static = c;
        } else {
            ctl->num_cells = c;
            if (nt > max_tiles) {
                ctl->err_tile_count = nt;
                nt = max_tiles;
            }
            ctl->num_tiles = nt;
            tile_pstart[nt] = n;
        }
    }
}

// lower bound over the compact cell list (the role of find_next_cell, neighborhood_search.rs:169-189)
__device__ __forceinline__ uint32_t cell_lower_bound(const uint32_t* __restrict__ cell_key, uint32_t ncells, uint32_t key) {
    uint32_t lo = 0, hi = ncells;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (cell_key[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// One CTA (128 threads) per tile, grid-stride: look up the 100 region cells in the dynamic and the static grid, order
// them by Morton key, assign shared-memory slots.
__global__ void __launch_bounds__(128)
    k_tile_tables(const uint32_t* __restrict__ tile_key, const uint32_t* __restrict__ tile_pstart, const uint32_t* __restrict__ cell_key,
                  const uint32_t* __restrict__ cell_start, const uint32_t* __restrict__ scell_key, const uint32_t* __restrict__ scell_start,
                  Control* ctl, TileHeader* __restrict__ hdr, TileCell* __restrict__ tdyn, TileCell* __restrict__ tstat,
                  uint8_t* __restrict__ rank_of_cell, uint32_t cap_dyn, uint32_t cap_stat) {
    __shared__ unsigned long long skey[YASPH_REGION_CELLS];
    __shared__ uint32_t gs_d[YASPH_REGION_CELLS], cn_d[YASPH_REGION_CELLS], gs_s[YASPH_REGION_CELLS], cn_s[YASPH_REGION_CELLS];
    __shared__ uint32_t by_d[YASPH_REGION_CELLS], by_s[YASPH_REGION_CELLS];
    __shared__ uint32_t own_rank;
    const uint32_t ntiles = ctl->num_tiles, ncells = ctl->num_cells, nscells = ctl->num_cells_static;
    const uint32_t r = threadIdx.x;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint32_t k0 = tile_key[t] << YASPH_TILE_SHIFT;
        const int x0 = (int)morton_x(k0), y0 = (int)morton_y(k0);
        uint32_t myrank = 0;
        if (r < YASPH_REGION_CELLS) {
            const int lx = (int)(r % YASPH_REGION_AXIS), ly = (int)(r / YASPH_REGION_AXIS);
            const int x = x0 + lx - 1, y = y0 + ly - 1;
            const bool valid = x >= 0 && x <= 65535 && y >= 0 && y <= 65535;
            uint32_t gd = 0, cd = 0, gs = 0, cs = 0;
            unsigned long long sk = (1ull << 32) | r;
            if (valid) {
                uint32_t key = morton_encode((uint32_t)x, (uint32_t)y);
                sk = key;
                uint32_t ci = cell_lower_bound(cell_key, ncells, key);
                if (ci < ncells && cell_key[ci] == key) {
                    gd = cell_start[ci];
                    cd = cell_start[ci + 1] - gd;
                }
                ci = cell_lower_bound(scell_key, nscells, key);
                if (ci < nscells && scell_key[ci] == key) {
                    gs = scell_start[ci];
                    cs = scell_start[ci + 1] - gs;
                }
            }
            skey[r] = sk;
            gs_d[r] = gd;
            cn_d[r] = cd;
            gs_s[r] = gs;
            cn_s[r] = cs;
        }
        __syncthreads();
        if (r < YASPH_REGION_CELLS) {
            const unsigned long long mine = skey[r];
            uint32_t rk = 0;
            for (int q = 0; q < YASPH_REGION_CELLS; ++q) rk += skey[q] < mine ? 1u : 0u;
            myrank = rk;
            by_d[rk] = cn_d[r];
            by_s[rk] = cn_s[r];
            rank_of_cell[(size_t)t * 128 + r] = (uint8_t)rk;
            if (r == YASPH_REGION_AXIS + 1) own_rank = rk;  // region cell (1,1) == first own cell
        }
        __syncthreads();
        if (r < YASPH_REGION_CELLS) {
            uint32_t sd = 0, ss = 0;
            for (uint32_t q = 0; q < myrank; ++q) {
                sd += by_d[q];
                ss += by_s[q];
            }
            // slot_start / count are packed in 16 bits each; totals above the capacity are flagged below and the
            // tile is then never consumed (the step reports YASPH_ERR_CAPACITY)
            tdyn[(size_t)t * YASPH_REGION_CELLS + myrank] = make_uint2(gs_d[r], (sd & 0xFFFFu) | (cn_d[r] << 16));
            tstat[(size_t)t * YASPH_REGION_CELLS + myrank] = make_uint2(gs_s[r], (ss & 0xFFFFu) | (cn_s[r] << 16));
            if (myrank == YASPH_REGION_CELLS - 1) {
                TileHeader h;
                h.pstart = tile_pstart[t];
                h.pcount = tile_pstart[t + 1] - h.pstart;
                h.dyn_total = sd + cn_d[r];
                h.stat_total = ss + cn_s[r];
                uint32_t lo = 0;
                for (uint32_t q = 0; q < own_rank; ++q) lo += by_d[q];
                h.own_lo = lo;
                h.pad0 = h.pad1 = h.pad2 = 0;
                hdr[t] = h;
                if (h.dyn_total > cap_dyn || h.stat_total > cap_stat) atomicMax(&ctl->err_tile_capacity, max(h.dyn_total, h.stat_total));
            }
        }
        __syncthreads();
    }
}

// ---- staging helpers (shared by the list build and every sweep) ----------------------------------------------------------
struct TileSmem {
    TileHeader hdr;
    TileCell dyn[YASPH_REGION_CELLS];
    TileCell stat[YASPH_REGION_CELLS];
    uint8_t rank[128];
};

__device__ __forceinline__ void load_tile_tables(TileSmem& ts, const TileTables& tt, uint32_t t) {
    for (uint32_t q = threadIdx.x; q < YASPH_REGION_CELLS; q += blockDim.x) {
        ts.dyn[q] = tt.dyn[(size_t)t * YASPH_REGION_CELLS + q];
        ts.stat[q] = tt.stat[(size_t)t * YASPH_REGION_CELLS + q];
    }
    for (uint32_t q = threadIdx.x; q < 128 / 4; q += blockDim.x)
        reinterpret_cast<uint32_t*>(ts.rank)[q] = reinterpret_cast<const uint32_t*>(tt.rank_of_cell + (size_t)t * 128)[q];
    if (threadIdx.x == 0) ts.hdr = tt.hdr[t];
}
// global index of staged slot s: the last table entry (rank order) whose slot_start <= s
__device__ __forceinline__ uint32_t slot_to_global(const TileCell* tab, uint32_t s) {
    uint32_t lo = 0, hi = YASPH_REGION_CELLS - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if ((tab[mid].y & 0xFFFFu) <= s)
            lo = mid;
        else
            hi = mid - 1;
    }
    return tab[lo].x + (s - (tab[lo].y & 0xFFFFu));
}
__device__ __forceinline__ uint32_t dyn_slot_to_global(const TileSmem& ts, uint32_t s) {
    uint32_t o = s - ts.hdr.own_lo;  // own particles are one contiguous copy
    if (o < ts.hdr.pcount) return ts.hdr.pstart + o;
    return slot_to_global(ts.dyn, s);
}
// 128-bit mask of the ranks of the 3x3 cells around the particle's cell (key & 63 = cell inside the tile)
__device__ __forceinline__ void neighbor_cell_mask(const TileSmem& ts, uint32_t key, unsigned long long& m0, unsigned long long& m1) {
    const uint32_t local = key & 63u;
    const uint32_t lx = morton_x(local) + 1, ly = morton_y(local) + 1;
    m0 = 0ull;
    m1 = 0ull;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            uint32_t rk = ts.rank[(ly + dy) * YASPH_REGION_AXIS + (lx + dx)];
            if (rk < 64)
                m0 |= 1ull << rk;
            else
                m1 |= 1ull << (rk - 64);
        }
}

// list storage: per tile, base = pstart * 64 entries; entry (k, tl) lives in 8-byte word (k/4)*pcount + tl, lane k%4
__device__ __forceinline__ size_t list_word_index(uint32_t pstart, uint32_t pcount, uint32_t kb, uint32_t tl) {
    return (size_t)pstart * (YASPH_MAXN / 4) + (size_t)kb * pcount + tl;
}
__device__ __forceinline__ unsigned long long pack_slot(unsigned long long w, uint32_t k, uint32_t slot) {
    const int sh = (int)(k & 3u) * 16;
    return (w & ~(0xFFFFull << sh)) | ((unsigned long long)slot << sh);
}
__device__ __forceinline__ uint32_t unpack_slot(unsigned long long w, uint32_t k) { return (uint32_t)(w >> ((k & 3u) * 16)) & 0xFFFFu; }

// One CTA per tile (grid-stride): NeighborLists::try_update (neighborhood_search.rs:312-397)
__global__ void __launch_bounds__(NB_THREADS)
    k_build_lists(TileTables tt, const float2* __restrict__ pos, const float2* __restrict__ bpos, const uint32_t* __restrict__ keys,
                  GridParams g, Control* ctl, unsigned long long* __restrict__ lists, uchar2* __restrict__ counts, uint32_t cap_dyn, uint32_t cap_stat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem& ts = *reinterpret_cast<TileSmem*>(smem_raw);
    float2* sdyn = reinterpret_cast<float2*>(smem_raw + sizeof(TileSmem));
    float2* sstat = sdyn + cap_dyn;
    __shared__ unsigned long long s_total;
    const uint32_t ntiles = ctl->num_tiles;
    unsigned long long my_total = 0;
    uint32_t my_capped = 0, my_dropped = 0;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        load_tile_tables(ts, tt, t);
        __syncthreads();
        const TileHeader h = ts.hdr;
        if (h.dyn_total <= cap_dyn && h.stat_total <= cap_stat) {
            for (uint32_t s = threadIdx.x; s < h.dyn_total; s += blockDim.x) sdyn[s] = pos[dyn_slot_to_global(ts, s)];
            for (uint32_t s = threadIdx.x; s < h.stat_total; s += blockDim.x) sstat[s] = bpos[slot_to_global(ts.stat, s)];
            __syncthreads();
            for (uint32_t tl = threadIdx.x; tl < h.pcount; tl += blockDim.x) {
                const uint32_t i = h.pstart + tl;
                const float2 q = sdyn[h.own_lo + tl];
                unsigned long long m0, m1;
                neighbor_cell_mask(ts, keys[i], m0, m1);
                uint32_t cd = 0;
                unsigned long long e = 0ull;  // four u16 slots, entry k in bits [16*(k&3), +16)
                bool full = false;
                // dynamic candidates, ascending slot == ascending sorted index (neighborhood_search.rs:353-366)
                for (int half = 0; half < 2 && !full; ++half) {
                    unsigned long long m = half ? m1 : m0;
                    while (m && !full) {
                        const int rk = __ffsll((long long)m) - 1 + half * 64;
                        m &= m - 1;
                        const TileCell c = ts.dyn[rk];
                        const uint32_t s0 = c.y & 0xFFFFu, s1 = s0 + (c.y >> 16);
                        for (uint32_t s = s0; s < s1; ++s) {
                            const float2 d = sdyn[s] - q;
                            const float d2 = d.x * d.x + d.y * d.y;
                            if (d2 <= g.radius_sq && d2 > YASPH_MIN_DISTANCE) {
                                e = pack_slot(e, cd, s);
                                ++cd;
                                if ((cd & 3) == 0) lists[list_word_index(h.pstart, h.pcount, (cd >> 2) - 1, tl)] = e;
                                if (cd == YASPH_MAXN) {
                                    full = true;
                                    ++my_capped;
                                    break;
                                }
                            }
                        }
                    }
                }
                uint32_t ct = cd;
                full = false;
                // static candidates (neighborhood_search.rs:367-381)
                for (int half = 0; half < 2 && !full; ++half) {
                    unsigned long long m = half ? m1 : m0;
                    while (m && !full) {
                        const int rk = __ffsll((long long)m) - 1 + half * 64;
                        m &= m - 1;
                        const TileCell c = ts.stat[rk];
                        const uint32_t s0 = c.y & 0xFFFFu, s1 = s0 + (c.y >> 16);
                        for (uint32_t s = s0; s < s1; ++s) {
                            const float2 d = sstat[s] - q;
                            const float d2 = d.x * d.x + d.y * d.y;
                            if (d2 <= g.radius_sq && d2 > YASPH_MIN_DISTANCE) {
                                if (ct >= YASPH_MAXN) {  // the reference indexes neighbor_set[64] here and panics (:373)
                                    ++my_dropped;
                                    full = true;
                                    break;
                                }
                                e = pack_slot(e, ct, s);
                                ++ct;
                                if ((ct & 3) == 0) lists[list_word_index(h.pstart, h.pcount, (ct >> 2) - 1, tl)] = e;
                                if (ct == YASPH_MAXN) {
                                    full = true;
                                    ++my_capped;
                                    break;
                                }
                            }
                        }
                    }
                }
                if (ct & 3) lists[list_word_index(h.pstart, h.pcount, ct >> 2, tl)] = e;
                counts[i] = make_uchar2((unsigned char)cd, (unsigned char)ct);
                my_total += ct;
            }
        }
        __syncthreads();
    }
    // statistics: one atomic per CTA
    if (threadIdx.x == 0) s_total = 0ull;
    __syncthreads();
    if (my_total) atomicAdd(&s_total, my_total);
    if (my_capped) atomicAdd(&ctl->capped, my_capped);
    if (my_dropped) atomicAdd(&ctl->dropped, my_dropped);
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd(&ctl->total_neighbors, s_total);
}

// Export to the reference's layout (neighborhood_search.rs:268-273,433-449): u16 counts + u32 global indices, stride 64.
__global__ void __launch_bounds__(NB_THREADS)
    k_export_lists(TileTables tt, const Control* __restrict__ ctl, const unsigned long long* __restrict__ lists, const uchar2* __restrict__ counts,
                   uint16_t* __restrict__ out_cd, uint16_t* __restrict__ out_ct, uint32_t* __restrict__ out_lists) {
    __shared__ TileSmem ts;
    const uint32_t ntiles = ctl->num_tiles;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        load_tile_tables(ts, tt, t);
        __syncthreads();
        const TileHeader h = ts.hdr;
        for (uint32_t tl = threadIdx.x; tl < h.pcount; tl += blockDim.x) {
            const uint32_t i = h.pstart + tl;
            const uchar2 c = counts[i];
            out_cd[i] = c.x;
            out_ct[i] = c.y;
            if (out_lists) {
                for (uint32_t k = 0; k < c.y; ++k) {
                    const uint32_t s = unpack_slot(lists[list_word_index(h.pstart, h.pcount, k >> 2, tl)], k);
                    out_lists[(size_t)i * YASPH_MAXN + k] = k < c.x ? dyn_slot_to_global(ts, s) : slot_to_global(ts.stat, s);
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace yasph
