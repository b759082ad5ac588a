// neighborhood.cuh -- Morton-sorted compact cell grid, cell tiles, per-tile staging tables and the neighbour-list build.
//
// Replaces CompactMortonCellGrid::update (src/sph/neighborhood_search.rs:90-166), the per-cell run search
// get_particle_runs_in_neighborbox (:191-259) and NeighborLists::try_update (:312-397).  Semantics kept from the
// reference: candidates of a particle are all particles in the 3x3 cell box of its cell, visited in ascending sorted
// index (== ascending Morton code of the cell; :191-259 produces exactly that order as <=5 index runs); a candidate is a
// neighbour iff d2 <= r2 && d2 > 1e-10 with d2 = fl(fl(dx*dx) + fl(dy*dy)) (:356-357); dynamic neighbours first, then
// static, at most 64 in total (:322).
//
// B200 design: particles are stored Morton-sorted, so an aligned 8x8 block of cells (a "tile") is one contiguous index
// range.  Every neighbour-dependent pass runs one CTA per tile, copies the tile's particles plus its 1-cell apron into
// shared memory in ascending-key order and addresses neighbours by 16-bit shared-memory slot.  Ascending slot ==
// ascending global index, so list order (and with it every floating-point sum order) is the reference's.  Lists are
// stored as u16 slots, 4 per 8-byte word, interleaved across the tile's particles so a warp reads them coalesced: 2 B per
// neighbour instead of the reference's 4 B + 8 B range record (neighborhood_search.rs:268-273,299).
//
// Per tile k_tile_tables produces
//   TileRuns   header + the maximal contiguous global index runs ("copy runs") that make up the staged dynamic and
//              static candidate arrays -- what the staging loops of every tile kernel binary-search (<= 37 entries)
//   cslot_d/s  per region cell (row-major in the 10x10 region): slot_start << 16 | count -- what the list build uses
//              to enumerate the 3x3 box of a cell
// and the maxima over all tiles (Control::max_*), which the host reads back once per neighbourhood update to size the
// shared memory of the tile kernels exactly (no worst-case capacity, so many CTAs fit on an SM).
#pragma once
#include "sort.cuh"

namespace yasph {

constexpr int TILE_AXIS = 1 << YASPH_TILE_LOG2;
constexpr int TILE_CELLS = TILE_AXIS * TILE_AXIS;
constexpr int TILE_SHIFT = 2 * YASPH_TILE_LOG2;  // key >> TILE_SHIFT == tile key
constexpr int REGION_AXIS = TILE_AXIS + 2;
constexpr int REGION_CELLS = REGION_AXIS * REGION_AXIS;
constexpr int APRON_CELLS = 4 * TILE_AXIS + 4;
constexpr int MAX_RUNS = (APRON_CELLS + 1 + 3) & ~3;  // every apron cell on its own + the own block, rounded up
constexpr uint32_t MAX_BLOCK_COORD = 65535u >> YASPH_TILE_LOG2;
static_assert(REGION_CELLS <= 128, "the tile-table warp handles the region in four rounds of 32 lanes");

struct TileHeader {       // 32 bytes
    uint32_t pstart;      // first particle of the tile (sorted index)
    uint32_t pcount;      // particles in the tile
    uint32_t dyn_total;   // staged dynamic candidates (tile + apron)
    uint32_t stat_total;  // staged boundary candidates
    uint32_t own_lo;      // slot of the tile's first own particle in the staged dynamic array
    uint32_t nruns_d, nruns_s;
    uint32_t pad;         // unused slots before | after << 8 the tile's own particles (see k_tile_tables step 4)
};
// The own particles of a tile sit at slots [own_lo, own_lo + pcount) with own_lo == pstart (mod 4) and are surrounded by
// pad slots such that the 4-element-aligned global range [pstart & ~3, roundup4(pstart + pcount)) maps onto the
// 4-slot-aligned range [own_lo - (pstart & 3), ...): every per-particle array can then be staged with ONE 16-byte aligned bulk
// copy, whose leading / trailing surplus elements land in the pad slots.
__host__ __device__ __forceinline__ uint32_t tile_pad_before(const TileHeader& h) { return h.pad & 0xFFu; }
__host__ __device__ __forceinline__ uint32_t tile_pad_after(const TileHeader& h) { return (h.pad >> 8) & 0xFFu; }
// the a-th staged dynamic candidate that is NOT one of the tile's own particles or a pad slot (a < tile_apron_count)
__host__ __device__ __forceinline__ uint32_t tile_apron_count(const TileHeader& h) { return h.dyn_total - h.pcount - tile_pad_before(h) - tile_pad_after(h); }
__host__ __device__ __forceinline__ uint32_t tile_apron_slot(const TileHeader& h, uint32_t a) {
    const uint32_t lo = h.own_lo - tile_pad_before(h);
    return a < lo ? a : a - lo + h.own_lo + h.pcount + tile_pad_after(h);
}
// copy run: .x = first global index, .y = first slot; unused entries carry .y = 0xFFFFFFFF
struct TileRuns {
    TileHeader hdr;
    uint2 rd[MAX_RUNS];
    uint2 rs[MAX_RUNS];
};
static_assert(sizeof(TileRuns) % 16 == 0, "TileRuns is copied in 16-byte pieces");

struct TileTables {
    const TileRuns* runs;     // [tile]
    const uint32_t* cslot_d;  // [tile][REGION_CELLS]
    const uint32_t* cslot_s;  // [tile][REGION_CELLS]
};

// ---- keys ---------------------------------------------------------------------------------------------------------
// All three key-generating kernels run KG_THREADS threads per block and feed the radix sort's global digit histograms
// (sort.cuh) while they have the key in a register.
// grid of a key-generating kernel: a few CTAs per SM, each walking chunks of KG_THREADS particles
inline uint32_t keygen_grid(uint32_t n, int num_sms) {
    const uint32_t chunks = (n + KG_THREADS - 1) / KG_THREADS, cap = (uint32_t)num_sms * 8u;
    return chunks < cap ? (chunks ? chunks : 1u) : cap;
}
// neighborhood_search.rs:111-114 (sequential in the reference)
__global__ void __launch_bounds__(KG_THREADS)
    k_keygen(const float2* __restrict__ pos, uint32_t first, uint32_t n, GridParams g, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx,
             uint32_t* __restrict__ sort_scratch, SlabParams sp, uint32_t classify_end) {
    // classify_end: particles from this index on are not classified (the ghost arrivals of a slab exchange, which lie outside the slab
    // by construction, behind the migrant arrivals, which must lie inside)
    __shared__ RadixHistSmem sh;
    radix_hist_init(sh);
    pdl_enter();
    for (uint32_t base = first + blockIdx.x * KG_THREADS; base < n; base += gridDim.x * KG_THREADS) {
        const uint32_t i = base + threadIdx.x;
        uint32_t key = 0;
        if (i < n) {
            key = position_to_cidx(g, pos[i]);
            if (i < classify_end) key = slab_classify(sp, i, key);
            keys[i] = key;
        }
        radix_hist_add(sh, key, i < n);
    }
    radix_hist_flush(sh, sort_scratch);
}
// dfsph.rs:502-509 (advect) fused with the key generation of the following re-sort
__global__ void __launch_bounds__(KG_THREADS)
    k_advect_keygen(float2* __restrict__ pos, const float2* __restrict__ vstar, uint32_t n, const Control* __restrict__ ctl, GridParams g,
                    uint32_t* __restrict__ keys, uint32_t* __restrict__ idx, uint32_t* __restrict__ sort_scratch, SlabParams sp, uint32_t only_if_converged) {
    __shared__ RadixHistSmem sh;
    radix_hist_init(sh);
    pdl_enter();
    // launched ahead of the density solver's read-back (dfsph_step): nothing may move unless that solve has finished
    if (only_if_converged && ctl->stop_iter[0] == 0xFFFFFFFFu) return;
    const float dt = ctl->dt;
    for (uint32_t base = blockIdx.x * KG_THREADS; base < n; base += gridDim.x * KG_THREADS) {
        const uint32_t i = base + threadIdx.x;
        uint32_t key = 0;
        if (i < n) {
            float2 p = pos[i] + vstar[i] * dt;
            pos[i] = p;
            key = slab_classify(sp, i, position_to_cidx(g, p));
            keys[i] = key;
        }
        radix_hist_add(sh, key, i < n);
    }
    radix_hist_flush(sh, sort_scratch);
}
// wscsph.rs:141-150 (leap frog 1) fused with key generation
__global__ void __launch_bounds__(KG_THREADS)
    k_kickdrift_keygen(float2* __restrict__ pos, float2* __restrict__ vel, const float2* __restrict__ acc, uint32_t n, const Control* __restrict__ ctl,
                       GridParams g, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx, uint32_t* __restrict__ sort_scratch, SlabParams sp) {
    __shared__ RadixHistSmem sh;
    radix_hist_init(sh);
    pdl_enter();
    const float dt = ctl->dt_prev;
    for (uint32_t base = blockIdx.x * KG_THREADS; base < n; base += gridDim.x * KG_THREADS) {
        const uint32_t i = base + threadIdx.x;
        uint32_t key = 0;
        if (i < n) {
            float2 v = vel[i] + 0.5f * dt * acc[i];
            float2 p = pos[i] + v * dt;
            vel[i] = v;
            pos[i] = p;
            key = slab_classify(sp, i, position_to_cidx(g, p));
            keys[i] = key;
        }
        radix_hist_add(sh, key, i < n);
    }
    radix_hist_flush(sh, sort_scratch);
}

// apply_sorting (neighborhood_search.rs:71-78): out[k] = in[perm[k]] for up to three float2 and three 4-byte arrays
struct GatherArgs {
    const float2* in2[3];
    float2* out2[3];
    const float* in1[3];
    float* out1[3];
    int n2, n1;
    // slab mode (null otherwise): ghost flags of the NEW structure from the sorted keys (a particle whose cell column lies outside
    // [col_lo, col_hi) is a ghost copy of a particle the left / right rank owns) and their counts (left | right << 32)
    const uint32_t* keys;
    uint8_t* ghost;
    unsigned long long* ghost_count;
    uint32_t col_lo, col_hi;
};
__global__ void k_gather(const uint32_t* __restrict__ perm, uint32_t n, GatherArgs a) {
    pdl_enter();
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    bool gl = false, gr = false;
    if (k < n) {
        uint32_t s = perm[k];
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q < a.n2) a.out2[q][k] = a.in2[q][s];
#pragma unroll
        for (int q = 0; q < 3; ++q)
            if (q < a.n1) a.out1[q][k] = a.in1[q][s];
        if (a.ghost) {
            const uint32_t col = compact_1by1(a.keys[k]);
            gl = col < a.col_lo;
            gr = col >= a.col_hi;
            a.ghost[k] = (gl || gr) ? 1 : 0;
        }
    }
    if (a.ghost) {
        const unsigned ml = __ballot_sync(0xffffffffu, gl), mr = __ballot_sync(0xffffffffu, gr);
        if ((ml | mr) && lane_id() == 0) atomicAdd(a.ghost_count, (unsigned long long)__popc(ml) | ((unsigned long long)__popc(mr) << 32));
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
        unsigned long long t = (i == 0 || (k >> TILE_SHIFT) != (p >> TILE_SHIFT)) ? 1ull : 0ull;
        return c | (t << 32);
    }
};
struct HeadCompactOut {
    const uint32_t* keys;
    uint32_t* cell_key;
    uint32_t* cell_start;
    uint32_t* tile_key;
    uint32_t* tile_pstart;  // may be null (static grid)
    uint32_t* tile_cstart;
    uint32_t max_tiles;
    __device__ __forceinline__ void operator()(uint32_t i, unsigned long long ex, unsigned long long v) const {
        const uint32_t c = (uint32_t)(ex & 0xFFFFFFFFull);
        if (v & 0xFFFFFFFFull) {
            cell_key[c] = keys[i];
            cell_start[c] = i;
        }
        if (v >> 32) {
            uint32_t t = (uint32_t)(ex >> 32);
            if (t < max_tiles) {
                tile_key[t] = keys[i] >> TILE_SHIFT;
                if (tile_pstart) tile_pstart[t] = i;
                tile_cstart[t] = c;  // a tile head is a cell head: c is the index of the tile's first cell
            }
        }
    }
};
// sentinel cell {first_particle = N, cidx = u32::MAX} (neighborhood_search.rs:161-164), tile sentinels and the counts, from the
// scan's grand total (cells | tiles << 32): called by the thread of the fused scan that holds it
struct FinishCells {
    uint32_t n;
    uint32_t *cell_key, *cell_start, *tile_pstart, *tile_cstart;
    uint32_t max_tiles;
    Control* ctl;
    int is_static;
    __device__ __forceinline__ void operator()(unsigned long long t) const {
        uint32_t c = (uint32_t)(t & 0xFFFFFFFFull), nt = (uint32_t)(t >> 32);
        cell_key[c] = 0xFFFFFFFFu;
        cell_start[c] = n;
        if (nt > max_tiles) {
            ctl->err_tile_count = nt;
            nt = max_tiles;
        }
        tile_cstart[nt] = c;
        if (is_static) {
            ctl->num_cells_static = c;
            ctl->num_tiles_static = nt;
        } else {
            ctl->num_cells = c;
            ctl->num_tiles = nt;
            tile_pstart[nt] = n;
            ctl->max_dyn_total = 0u;
            ctl->max_stat_total = 0u;
            ctl->max_pcount = 0u;
            ctl->max_nk = 0u;
            ctl->total_neighbors = 0ull;  // the list build's statistics
            ctl->capped = 0u;
            ctl->dropped = 0u;
        }
    }
};
// the same for an empty particle set
__global__ void k_finish_cells(FinishCells f) {
    if (threadIdx.x == 0 && blockIdx.x == 0) f(0ull);
}

// first index in [lo, hi) whose key is >= `key` (the role of find_next_cell, neighborhood_search.rs:169-189)
__device__ __forceinline__ uint32_t key_lower_bound(const uint32_t* __restrict__ keys, uint32_t lo, uint32_t hi, uint32_t key) {
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] < key)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// region cell (rx, ry) in [0, REGION_AXIS)^2 -> which of the 3x3 tile-sized blocks it lies in, and its position among
// the region cells of that block in ascending Morton order
__device__ __forceinline__ uint32_t region_block(uint32_t rx, uint32_t ry) {
    const uint32_t bx = rx == 0 ? 0u : (rx == REGION_AXIS - 1 ? 2u : 1u);
    const uint32_t by = ry == 0 ? 0u : (ry == REGION_AXIS - 1 ? 2u : 1u);
    return by * 3u + bx;
}
__device__ __forceinline__ uint32_t region_order_in_block(uint32_t rx, uint32_t ry, uint32_t b) {
    const uint32_t bx = b % 3u, by = b / 3u;
    if (bx == 1u && by == 1u) return morton_encode(rx - 1u, ry - 1u);  // own block: local Morton code
    if (by == 1u) return ry - 1u;                                      // west / east column: one local x, ascending y
    if (bx == 1u) return rx - 1u;                                      // south / north row
    return 0u;                                                         // corner
}
__device__ __forceinline__ uint32_t region_cells_in_block(uint32_t b) {
    const uint32_t bx = b % 3u, by = b / 3u;
    return (bx == 1u ? (uint32_t)TILE_AXIS : 1u) * (by == 1u ? (uint32_t)TILE_AXIS : 1u);
}

// One warp per tile: look the region cells up in the dynamic and the static grid, order them by Morton key, assign
// shared-memory slots, merge them into copy runs.
#ifndef YASPH_TT_WARPS
#define YASPH_TT_WARPS 4
#endif
constexpr int TT_WARPS = YASPH_TT_WARPS;
struct TTScratch {
    uint32_t cnt[2][REGION_CELLS];   // [dynamic | static][region cell, row-major]
    uint32_t gs[2][REGION_CELLS];    // first global index
    uint32_t slot[2][REGION_CELLS];  // first slot
    uint32_t blk_lo[2][9], blk_hi[2][9];  // the block's range in the cell list
    uint32_t blk_seq[9];             // position of the block's first region cell in slot order
    uint8_t seq[128];                // region cells in slot order
};
struct TileTableArgs {
    const uint32_t *tile_key, *tile_pstart, *tile_cstart, *cell_key, *cell_start;  // dynamic grid
    const uint32_t *stile_key, *stile_cstart, *scell_key, *scell_start;            // static grid
    TileRuns* truns;
    uint32_t *cslot_d, *cslot_s;
};
__global__ void __launch_bounds__(TT_WARPS * 32) k_tile_tables(TileTableArgs a, Control* ctl) {
    __shared__ TTScratch scratch[TT_WARPS];
    TTScratch& S = scratch[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const unsigned lt = lanemask_lt();
    pdl_enter();
    const uint32_t ntiles = ctl->num_tiles, nstiles = ctl->num_tiles_static;
    uint32_t wmax_d = 0, wmax_s = 0, wmax_p = 0;
    for (uint32_t t = blockIdx.x * TT_WARPS + (threadIdx.x >> 5); t < ntiles; t += gridDim.x * TT_WARPS) {
        const uint32_t tkey = a.tile_key[t];
        const uint32_t k0 = tkey << TILE_SHIFT;
        const int x0 = (int)morton_x(k0), y0 = (int)morton_y(k0);
        const uint32_t pstart = a.tile_pstart[t], pend = a.tile_pstart[t + 1];
        for (uint32_t r = lane; r < REGION_CELLS; r += 32) {
            S.cnt[0][r] = 0;
            S.cnt[1][r] = 0;
            S.gs[0][r] = 0;
            S.gs[1][r] = 0;
        }
        // 1. the nine blocks: key, rank, cell ranges in both grids
        {
            uint32_t bkey = 0xFFFFFFFFu;
            if (lane < 9) {
                const int bx = (x0 >> YASPH_TILE_LOG2) - 1 + (int)(lane % 3u), by = (y0 >> YASPH_TILE_LOG2) - 1 + (int)(lane / 3u);
                const bool valid = bx >= 0 && bx <= (int)MAX_BLOCK_COORD && by >= 0 && by <= (int)MAX_BLOCK_COORD;
                if (valid) bkey = morton_encode((uint32_t)bx, (uint32_t)by);
                uint32_t lo = 0, hi = 0, slo = 0, shi = 0;
                if (lane == 4) {
                    lo = a.tile_cstart[t];
                    hi = a.tile_cstart[t + 1];
                } else if (valid) {
                    const uint32_t q = key_lower_bound(a.tile_key, 0, ntiles, bkey);
                    if (q < ntiles && a.tile_key[q] == bkey) {
                        lo = a.tile_cstart[q];
                        hi = a.tile_cstart[q + 1];
                    }
                }
                if (valid && nstiles) {
                    const uint32_t q = key_lower_bound(a.stile_key, 0, nstiles, bkey);
                    if (q < nstiles && a.stile_key[q] == bkey) {
                        slo = a.stile_cstart[q];
                        shi = a.stile_cstart[q + 1];
                    }
                }
                S.blk_lo[0][lane] = lo;
                S.blk_hi[0][lane] = hi;
                S.blk_lo[1][lane] = slo;
                S.blk_hi[1][lane] = shi;
            }
            // position of each block's first region cell in slot order: blocks ascend by key (invalid ones last, ties by lane)
            uint32_t seq0 = 0;
#pragma unroll
            for (uint32_t j = 0; j < 9; ++j) {
                const uint32_t kj = __shfl_sync(0xffffffffu, bkey, j);
                if (kj < bkey || (kj == bkey && j < lane)) seq0 += region_cells_in_block(j);
            }
            if (lane < 9) S.blk_seq[lane] = seq0;
        }
        __syncwarp();
        // most tiles have no boundary cell anywhere in their 3x3 blocks: the static half of steps 2 and 4 is skipped for them
        const bool any_static = __any_sync(0xffffffffu, lane < 9 && S.blk_lo[1][lane] < S.blk_hi[1][lane]);
        // 2. cell lookups.  Own block (both grids): walk the block's cells and scatter them; apron: search the block's range.
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            if (which == 1 && !any_static) break;
            const uint32_t* ckey = which ? a.scell_key : a.cell_key;
            const uint32_t* cstart = which ? a.scell_start : a.cell_start;
            const uint32_t olo = S.blk_lo[which][4], ohi = S.blk_hi[which][4];
            for (uint32_t c = olo + lane; c < ohi; c += 32) {
                const uint32_t key = ckey[c];
                const uint32_t r = (morton_y(key & (TILE_CELLS - 1)) + 1u) * REGION_AXIS + morton_x(key & (TILE_CELLS - 1)) + 1u;
                const uint32_t g = cstart[c];
                S.gs[which][r] = g;
                S.cnt[which][r] = cstart[c + 1] - g;
            }
            for (uint32_t ap = lane; ap < (uint32_t)APRON_CELLS; ap += 32) {
                uint32_t rx, ry;
                if (ap < (uint32_t)REGION_AXIS) {
                    rx = ap;
                    ry = 0;
                } else if (ap < 2u * REGION_AXIS) {
                    rx = ap - REGION_AXIS;
                    ry = REGION_AXIS - 1;
                } else if (ap < 2u * REGION_AXIS + TILE_AXIS) {
                    rx = 0;
                    ry = ap - 2u * REGION_AXIS + 1u;
                } else {
                    rx = REGION_AXIS - 1;
                    ry = ap - 2u * REGION_AXIS - TILE_AXIS + 1u;
                }
                const uint32_t b = region_block(rx, ry);
                const uint32_t lo = S.blk_lo[which][b], hi = S.blk_hi[which][b];
                if (lo < hi) {  // the block exists, so the cell coordinates are inside the domain
                    const uint32_t key = morton_encode((uint32_t)(x0 - 1 + (int)rx), (uint32_t)(y0 - 1 + (int)ry));
                    const uint32_t c = key_lower_bound(ckey, lo, hi, key);
                    if (c < hi && ckey[c] == key) {
                        const uint32_t g = cstart[c];
                        S.gs[which][ry * REGION_AXIS + rx] = g;
                        S.cnt[which][ry * REGION_AXIS + rx] = cstart[c + 1] - g;
                    }
                }
            }
        }
        // 3. slot order of the region cells
        for (uint32_t r = lane; r < REGION_CELLS; r += 32) {
            const uint32_t rx = r % REGION_AXIS, ry = r / REGION_AXIS;
            const uint32_t b = region_block(rx, ry);
            S.seq[S.blk_seq[b] + region_order_in_block(rx, ry, b)] = (uint8_t)r;
        }
        __syncwarp();
        // 4. walk the cells in slot order: slot starts (exclusive prefix of the counts) and copy runs
        uint32_t totals[2], nruns[2];
        TileRuns* out = a.truns + t;
        // pad slots around the own block (dynamic candidates only), see TileHeader
        const uint32_t ownB = S.blk_seq[4], ownE = ownB + TILE_CELLS;  // the own block's region cells in slot order: [ownB, ownE)
        uint32_t slotB = 0;
        for (uint32_t q = lane; q < ownB; q += 32) slotB += S.cnt[0][S.seq[q]];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) slotB += __shfl_xor_sync(0xffffffffu, slotB, o);
        const uint32_t pad_before = ((slotB + 3u) & ~3u) - slotB + (pstart & 3u);
        const uint32_t own_hi = slotB + pad_before + (pend - pstart);
        const uint32_t pad_after = ((own_hi + 3u) & ~3u) - own_hi;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            uint2* runs = which ? out->rs : out->rd;
            if (which == 1 && !any_static) {  // no boundary candidates: empty slot table, no runs
                for (uint32_t r = lane; r < REGION_CELLS; r += 32) S.slot[1][r] = 0u;
                totals[1] = 0u;
                nruns[1] = 0u;
                for (uint32_t q = lane; q < (uint32_t)MAX_RUNS; q += 32) runs[q] = make_uint2(0u, 0xFFFFFFFFu);
                break;
            }
            uint32_t carry = 0, carry_runs = 0, carry_end = 0, carry_q = 0;
            bool carry_valid = false;
            for (uint32_t q0 = 0; q0 < REGION_CELLS; q0 += 32) {
                const uint32_t q = q0 + lane;
                const bool in = q < REGION_CELLS;
                const uint32_t r = in ? S.seq[q] : 0u;
                const uint32_t cn = in ? S.cnt[which][r] : 0u;
                const uint32_t g = in ? S.gs[which][r] : 0u;
                const uint32_t ex = which == 0 ? (q == ownB ? pad_before : 0u) + (q == ownE ? pad_after : 0u) : 0u;  // pads precede the cell
                uint32_t inc = cn + ex;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= (uint32_t)o) inc += u;
                }
                const uint32_t slot = carry + inc - cn;
                if (in) S.slot[which][r] = slot;
                const unsigned ne = __ballot_sync(0xffffffffu, cn > 0);
                const unsigned before = ne & lt;
                const int pl = before ? 31 - __clz((int)before) : 0;
                const uint32_t pend_g = __shfl_sync(0xffffffffu, g + cn, pl);
                const bool has_prev = before ? true : carry_valid;
                const uint32_t prev_end = before ? pend_g : carry_end;
                const uint32_t prev_q = before ? q0 + (uint32_t)pl : carry_q;
                // a copy run never spans the pad slots on either side of the own block
                const bool crosses = which == 0 && ((prev_q < ownB && q >= ownB) || (prev_q < ownE && q >= ownE));
                const bool head = cn > 0 && !(has_prev && prev_end == g && !crosses);
                const unsigned hm = __ballot_sync(0xffffffffu, head);
                if (head) {
                    const uint32_t ri = carry_runs + (uint32_t)__popc(hm & lt);
                    if (ri < (uint32_t)MAX_RUNS) runs[ri] = make_uint2(g, slot);
                }
                carry += __shfl_sync(0xffffffffu, inc, 31);
                carry_runs += (uint32_t)__popc(hm);
                if (ne) {
                    const int ll = 31 - __clz((int)ne);
                    carry_end = __shfl_sync(0xffffffffu, g + cn, ll);
                    carry_q = q0 + (uint32_t)ll;
                    carry_valid = true;
                }
            }
            if (which == 0 && ownE >= (uint32_t)REGION_CELLS) carry += pad_after;  // the own block is last in slot order
            totals[which] = carry;
            nruns[which] = carry_runs < (uint32_t)MAX_RUNS ? carry_runs : (uint32_t)MAX_RUNS;
            for (uint32_t q = nruns[which] + lane; q < (uint32_t)MAX_RUNS; q += 32) runs[q] = make_uint2(0u, 0xFFFFFFFFu);
        }
        __syncwarp();
        // 5. header, per-cell slot tables, maxima
        if (lane == 0) {
            TileHeader h;
            h.pstart = pstart;
            h.pcount = pend - pstart;
            h.dyn_total = totals[0];
            h.stat_total = totals[1];
            h.own_lo = S.slot[0][REGION_AXIS + 1];  // region cell (1,1) == the tile's first cell
            h.nruns_d = nruns[0];
            h.nruns_s = nruns[1];
            h.pad = pad_before | (pad_after << 8);
            out->hdr = h;
            if (totals[0] > 0xFFFFu || totals[1] > 0xFFFFu) atomicMax(&ctl->err_tile_capacity, max(totals[0], totals[1]));
        }
        for (uint32_t r = lane; r < REGION_CELLS; r += 32) {
            a.cslot_d[(size_t)t * REGION_CELLS + r] = (S.slot[0][r] << 16) | (S.cnt[0][r] & 0xFFFFu);
            a.cslot_s[(size_t)t * REGION_CELLS + r] = (S.slot[1][r] << 16) | (S.cnt[1][r] & 0xFFFFu);
        }
        wmax_d = max(wmax_d, totals[0]);
        wmax_s = max(wmax_s, totals[1]);
        wmax_p = max(wmax_p, pend - pstart);
        __syncwarp();
    }
    if (lane == 0) {  // same-address atomics serialise in L2: only issue the ones that can still raise the maximum
        if (wmax_d > *(volatile unsigned int*)&ctl->max_dyn_total) atomicMax(&ctl->max_dyn_total, wmax_d);
        if (wmax_s > *(volatile unsigned int*)&ctl->max_stat_total) atomicMax(&ctl->max_stat_total, wmax_s);
        if (wmax_p > *(volatile unsigned int*)&ctl->max_pcount) atomicMax(&ctl->max_pcount, wmax_p);
    }
}

// ---- helpers shared by the tile kernels --------------------------------------------------------------------------------
__device__ __forceinline__ void load_tile_runs(TileRuns& dst, const TileRuns* __restrict__ src) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(&dst);
    for (uint32_t q = threadIdx.x; q < sizeof(TileRuns) / 16; q += blockDim.x) d[q] = s[q];
}
// global index of staged slot s: the last copy run whose first slot is <= s
__device__ __forceinline__ uint32_t run_slot_to_global(const uint2* runs, uint32_t s) {
    uint32_t lo = 0, hi = MAX_RUNS - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (runs[mid].y <= s)
            lo = mid;
        else
            hi = mid - 1;
    }
    return runs[lo].x + (s - runs[lo].y);
}
__device__ __forceinline__ uint32_t dyn_slot_to_global(const TileRuns& tr, uint32_t s) {
    uint32_t o = s - tr.hdr.own_lo;  // own particles are one contiguous copy
    if (o < tr.hdr.pcount) return tr.hdr.pstart + o;
    return run_slot_to_global(tr.rd, s);
}

// list storage: per tile, base = pstart * LIST_WORDS words; word kb of particle tl lives at kb * pcount + tl (a warp reads 256
// contiguous bytes).  A word holds four u16 slots.  Dynamic neighbours fill words [0, ceil(cd / 4)), the last one PADDED with the
// particle's own slot (a pair of a particle with itself contributes an exact zero to every pass but the density sum, so the
// sweeps need no per-entry validity test); static neighbours follow from the next word on.
constexpr int LIST_WORDS = YASPH_MAXN / 4 + 2;  // ceil(cd / 4) + ceil(cs / 4) <= 17 for cd + cs <= 64; even, so a tile's block is 16-byte aligned (bulk copies)
__device__ __forceinline__ size_t list_word_index(uint32_t pstart, uint32_t pcount, uint32_t kb, uint32_t tl) {
    return (size_t)pstart * LIST_WORDS + (size_t)kb * pcount + tl;
}
__device__ __forceinline__ uint32_t unpack_slot(unsigned long long w, uint32_t k) { return (uint32_t)(w >> ((k & 3u) * 16)) & 0xFFFFu; }

// compare-exchange for the 9-element sorting network below
__device__ __forceinline__ void cswap(uint32_t& a, uint32_t& b) {
    const uint32_t lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

// ---- cp.async staging (shared with sweeps.cuh) ---------------------------------------------------------------------------
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async.ca copies 4, 8 or 16 bytes");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int TILE_THREADS = 256;  // CTA size of every tile kernel
// the copy-run table of a tile travels global -> registers -> shared memory two tiles ahead of its use
struct RunsPrefetch {
    uint4 v;
    __device__ __forceinline__ void load(const TileRuns* __restrict__ src, bool valid) {
        static_assert(sizeof(TileRuns) / 16 <= TILE_THREADS, "one uint4 per thread");
        if (valid && threadIdx.x < sizeof(TileRuns) / 16) v = reinterpret_cast<const uint4*>(src)[threadIdx.x];
    }
    __device__ __forceinline__ void store(TileRuns& dst, bool valid) const {
        if (valid && threadIdx.x < sizeof(TileRuns) / 16) reinterpret_cast<uint4*>(&dst)[threadIdx.x] = v;
    }
};

// ---- neighbour lists: NeighborLists::try_update (neighborhood_search.rs:312-397) ---------------------------------------
// Persistent CTAs over tiles, staging double-buffered with cp.async like the sweeps (sweeps.cuh).  Per tile: one thread
// per own cell turns the 3x3 box of its cell into <= 9 slot runs in ascending slot order (the role of
// get_particle_runs_in_neighborbox, :191-259); then one thread per particle walks its cell's candidates in ONE flat loop
// (all lanes of a warp iterate about the same number of times, whatever their cells' run structure) whose body is
// branch-free: the candidate's slot is stored to the thread's column of a shared-memory list at row min(count, 64) and the
// count advances by the hit predicate, so the first 64 hits survive exactly as the reference's early exit leaves them.
// The column is then packed into 8-byte words of four u16 slots.
#ifndef YASPH_NB_THREADS
#define YASPH_NB_THREADS 256
#endif
constexpr int NB_THREADS = YASPH_NB_THREADS;  // CTA size of the list build (a multiple of 32, >= 2 * TILE_CELLS)
static_assert(NB_THREADS % 32 == 0 && NB_THREADS >= 2 * TILE_CELLS && sizeof(TileRuns) / 16 <= NB_THREADS, "list-build CTA size");
constexpr int NB_ROWS = YASPH_MAXN + 2;  // rows 64 and 65 absorb everything past the cap (65: "and then some", see list_scan_candidates)
constexpr int NB_COL_WORDS = NB_ROWS / 2;  // 33 words of two 16-bit rows
static_assert(NB_ROWS % 2 == 0 && NB_COL_WORDS % 2 == 1, "odd word stride per thread");
struct ListSmem {
    TileRuns runs[3];
    uint32_t cs[2][2][REGION_CELLS];  // [buffer][dynamic | static][region cell]: slot_start << 16 | count
    uint32_t crun[2][TILE_CELLS][9];  // per own cell: slot_start << 16 | count, ascending, merged
    uint32_t ncand[2][TILE_CELLS];    // per own cell: total candidates
    unsigned long long wtotal[NB_THREADS / 32];
    uint32_t nk_max;
    // hit columns, thread-major: thread t's rows are the NB_COL_WORDS words from sl[t][0] on.  The odd word stride spreads the lanes of
    // a warp over all banks whatever rows they are at, and the packing below reads two entries per load.
    uint32_t sl[NB_THREADS][NB_COL_WORDS];
};
inline size_t list_smem_bytes(uint32_t cap_dyn, uint32_t cap_stat) { return sizeof(ListSmem) + 2 * ((size_t)cap_dyn + cap_stat) * sizeof(float2); }

// Apron index table: the list build, which resolves every apron slot of a tile to its global particle index anyway, records
// the first APRON_TABLE of them per tile; the sweeps' producer warps (five sweeps per step) then stage the apron without
// walking the copy-run table again.  Slots beyond the table (very full aprons) fall back to the run look-up.
constexpr uint32_t APRON_TABLE = 192;
struct ListArgs {
    TileTables tt;
    const float2* pos;
    const float2* bpos;
    const uint32_t* keys;
    GridParams g;
    Control* ctl;
    unsigned long long* lists;
    uint32_t* counts;   // count_dynamic | count_total << 8 of particle i
    uint32_t* tile_nk;  // [tile] most list words of any particle of the tile
    uint32_t cap_dyn, cap_stat;
    uint32_t* apron_idx;  // [tile][APRON_TABLE]
    uint32_t unstaged;    // 1: the launch of k_build_lists<true>
#ifdef YASPH_LIST_TIMING
    unsigned long long* dbg;  // [8] cycle counters per phase (profiling builds only)
#endif
};
__device__ __forceinline__ void list_issue_stage(const ListArgs& a, uint32_t t, const TileRuns& tr, float2* sdyn, float2* sstat, uint32_t (*cs)[REGION_CELLS]) {
    const TileHeader& h = tr.hdr;
    for (uint32_t q = threadIdx.x; q < 2 * REGION_CELLS; q += NB_THREADS)
        cp_async<4>(&cs[q / REGION_CELLS][q % REGION_CELLS], (q < REGION_CELLS ? a.tt.cslot_d : a.tt.cslot_s) + (size_t)t * REGION_CELLS + q % REGION_CELLS);
    // a tile that outgrows the staging capacity is not staged: its candidates are read from global memory (k_build_lists)
    const bool fits = !a.unstaged && h.dyn_total <= a.cap_dyn && h.stat_total <= a.cap_stat;
    if (a.unstaged) return;  // the second launch (tiles too large to stage) needs the cell tables only
    if (fits)
        for (uint32_t s = threadIdx.x; s < h.pcount; s += NB_THREADS) cp_async<8>(&sdyn[h.own_lo + s], &a.pos[h.pstart + s]);
    for (uint32_t q = threadIdx.x, na = tile_apron_count(h); q < na && (fits || q < APRON_TABLE); q += NB_THREADS) {
        const uint32_t s = tile_apron_slot(h, q);
        const uint32_t g = run_slot_to_global(tr.rd, s);
        if (fits) cp_async<8>(&sdyn[s], &a.pos[g]);
        if (q < APRON_TABLE) a.apron_idx[(size_t)t * APRON_TABLE + q] = g;  // the sweeps may stage this tile even when this launch does not
    }
    if (fits)
        for (uint32_t s = threadIdx.x; s < h.stat_total; s += NB_THREADS) cp_async<8>(&sstat[s], &a.bpos[run_slot_to_global(tr.rs, s)]);
}
// Shared-memory accesses of the candidate scan by 32-bit shared address.  The load is a plain asm (no memory clobber): the staged
// positions are not written while a tile is scanned, so the compiler may schedule it across the hit-column stores, which it could
// not do with an ordinary load (possible alias), and the address is ONE LEA instead of a window-base computation per candidate.
__device__ __forceinline__ float2 lds_f2_saddr(uint32_t saddr) {
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32_saddr(uint32_t saddr) {  // the run table of a cell: written before the tile's barrier
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts_u16_saddr(uint32_t saddr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"((unsigned short)v) : "memory");
}
constexpr uint32_t NB_ROW_BYTES = sizeof(uint16_t);  // rows of a thread's hit column are adjacent
// candidates of one kind (dynamic / static) for one particle.  cand: shared address of the staged positions; wp: shared address of
// the thread's hit-column entry of the current row (row == hits so far); returns the advanced wp.  Every candidate's slot is stored
// at wp, which advances by one row on a hit and is clamped at row YASPH_MAXN + 1: branch-free, the first 64 hits survive, and the
// caller still sees whether there were exactly 64 or more (row 64 vs 65 when it starts a scan at row <= 64).
__device__ __forceinline__ uint32_t list_scan_candidates(uint32_t cand, const uint32_t* __restrict__ cruns, uint32_t ncand, float2 q, float radius_sq, uint32_t wp,
                                                         uint32_t wp_max) {
    // sr: slot << 16 | candidates left in the run -- the format of a run entry, so a new run is one load, and "next slot, one
    // candidate less" is one addition
    uint32_t rp = (uint32_t)__cvta_generic_to_shared(cruns), sr = 0;
    for (uint32_t j = 0; j < ncand; ++j) {
        if ((sr & 0xFFFFu) == 0u) {
            sr = lds_u32_saddr(rp);
            rp += 4u;
        }
        const uint32_t s = sr >> 16;
        const float2 d = lds_f2_saddr(cand + s * 8u) - q;
        const float d2 = mag2(d);  // fl(fl(dx * dx) + fl(dy * dy)), neighborhood_search.rs:356
        sts_u16_saddr(wp, s);
        wp = min(wp + ((d2 <= radius_sq && d2 > YASPH_MIN_DISTANCE) ? NB_ROW_BYTES : 0u), wp_max);
        sr += 0xFFFFu;  // slot + 1, left - 1
    }
    return wp;
}
// The same for a tile that is too large to stage: every candidate's position comes from global memory through the tile's copy runs
// (slot -> global index).  Slow and correct: the reference accepts any particle density (neighborhood_search.rs:353-381).
template <bool STATIC>
__device__ __forceinline__ uint32_t list_scan_candidates_global(const float2* __restrict__ gpos, const TileRuns* tr, const uint32_t* __restrict__ cruns, uint32_t ncand,
                                                             float2 q, float radius_sq, uint16_t* col, uint32_t c) {
    uint32_t r = 0, rem = 0, s = 0;
    for (uint32_t j = 0; j < ncand; ++j) {
        if (rem == 0) {
            const uint32_t run = cruns[r++];
            s = run >> 16;
            rem = run & 0xFFFFu;
        }
        const float2 d = gpos[STATIC ? run_slot_to_global(tr->rs, s) : dyn_slot_to_global(*tr, s)] - q;
        const float d2 = mag2(d);
        col[min(c, (uint32_t)YASPH_MAXN)] = (uint16_t)s;
        c += (d2 <= radius_sq && d2 > YASPH_MIN_DISTANCE) ? 1u : 0u;
        ++s;
        --rem;
    }
    return c;
}
// UNSTAGED == false: the tiles that fit the staging capacity (all of them, normally).  UNSTAGED == true: launched after it when
// some tiles do not fit -- the same kernel working on exactly those tiles, from global memory.
template <bool UNSTAGED>
__global__ void __launch_bounds__(NB_THREADS) k_build_lists(ListArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ListSmem& S = *reinterpret_cast<ListSmem*>(smem_raw);
    float2* const sdyn[2] = {reinterpret_cast<float2*>(smem_raw + sizeof(ListSmem)),
                             reinterpret_cast<float2*>(smem_raw + sizeof(ListSmem)) + a.cap_dyn + a.cap_stat};
    pdl_enter();
    const uint32_t ntiles = a.ctl->num_tiles;
    const uint32_t tid = threadIdx.x, G = gridDim.x;
    unsigned long long my_total = 0;
    uint32_t my_capped = 0, my_dropped = 0, cta_nk = 0;
    {
        const uint32_t t0 = blockIdx.x, t1 = blockIdx.x + G;
        if (t0 < ntiles) load_tile_runs(S.runs[0], a.tt.runs + t0);
        if (t1 < ntiles) load_tile_runs(S.runs[1], a.tt.runs + t1);
        __syncthreads();
        if (t0 < ntiles) list_issue_stage(a, t0, S.runs[0], sdyn[0], sdyn[0] + a.cap_dyn, S.cs[0]);
        cp_async_commit();
    }
    uint32_t k = 0;
#ifdef YASPH_LIST_TIMING
    long long tph[6] = {0, 0, 0, 0, 0, 0}, tz = clock64(), t_begin = tz;
#define LT_MARK(i) { long long now_ = clock64(); tph[i] += now_ - tz; tz = now_; }
#else
#define LT_MARK(i)
#endif
    for (uint32_t t = blockIdx.x; t < ntiles; t += G, ++k) {
        const uint32_t b = k & 1u;
        const float2* cdyn = sdyn[b];
        const uint32_t cdyn_s = (uint32_t)__cvta_generic_to_shared(cdyn), cstat_s = cdyn_s + a.cap_dyn * (uint32_t)sizeof(float2);  // shared addresses
        const TileRuns& tr = S.runs[k % 3u];
        RunsPrefetch pre;
        const bool have2 = t + 2 * G < ntiles;
        pre.load(a.tt.runs + t + 2 * G, have2);
        cp_async_wait_all();
        __syncthreads();  // this tile's copies have landed; everybody is done with the previous tile
        LT_MARK(0)  // waiting for the staged tile
        if (t + G < ntiles) list_issue_stage(a, t + G, S.runs[(k + 1) % 3u], sdyn[b ^ 1u], sdyn[b ^ 1u] + a.cap_dyn, S.cs[b ^ 1u]);
        cp_async_commit();
        LT_MARK(1)  // issuing the next tile's copies
        const TileHeader h = tr.hdr;
        const bool fits = h.dyn_total <= a.cap_dyn && h.stat_total <= a.cap_stat;
        const bool mine = fits != UNSTAGED;  // the tiles of this launch
        if (tid < 2 * TILE_CELLS) {
            const uint32_t which = tid / TILE_CELLS, lc = tid % TILE_CELLS;
            const uint32_t lx = morton_x(lc) + 1u, ly = morton_y(lc) + 1u;
            uint32_t v[9];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const uint32_t e = S.cs[b][which][(ly + dy - 1) * REGION_AXIS + lx + dx - 1];
                    v[dy * 3 + dx] = (e & 0xFFFFu) ? e : 0xFFFFFFFFu;  // empty cells sort last
                }
            // 25-exchange sorting network for 9 keys (slot_start in the high half: ascending slot order)
            cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
            cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
            cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
            cswap(v[0], v[3]); cswap(v[3], v[6]); cswap(v[0], v[3]);
            cswap(v[1], v[4]); cswap(v[4], v[7]); cswap(v[1], v[4]);
            cswap(v[2], v[5]); cswap(v[5], v[8]); cswap(v[2], v[5]);
            cswap(v[1], v[3]); cswap(v[5], v[7]); cswap(v[2], v[6]);
            cswap(v[4], v[6]); cswap(v[2], v[4]); cswap(v[2], v[3]);
            cswap(v[5], v[6]);
            uint32_t nr = 0, cur = 0xFFFFFFFFu, tot = 0;
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                if (v[q] != 0xFFFFFFFFu) {
                    tot += v[q] & 0xFFFFu;
                    if (cur != 0xFFFFFFFFu && (cur >> 16) + (cur & 0xFFFFu) == (v[q] >> 16)) {
                        cur += v[q] & 0xFFFFu;  // contiguous slots: extend (the sum stays <= 65535, checked by k_tile_tables)
                    } else {
                        if (cur != 0xFFFFFFFFu) S.crun[which][lc][nr++] = cur;
                        cur = v[q];
                    }
                }
            }
            if (cur != 0xFFFFFFFFu) S.crun[which][lc][nr++] = cur;
            S.ncand[which][lc] = tot;
        }
        if (tid == 0) S.nk_max = 0u;
        __syncthreads();
        LT_MARK(2)  // cell phase (incl. its barrier)
        uint32_t my_nk = 0;
        if (mine) {
            for (uint32_t tl = tid; tl < h.pcount; tl += NB_THREADS) {
                const uint32_t i = h.pstart + tl;
                const float2 q = !UNSTAGED ? cdyn[h.own_lo + tl] : a.pos[i];
                const uint32_t lc = a.keys[i] & (TILE_CELLS - 1);
                uint16_t* col = reinterpret_cast<uint16_t*>(&S.sl[tid][0]);
                // hits_d / c: hit counts after the dynamic / static scan; the staged scan reports them clamped at 65 (enough for
                // everything below: 64 = the cap is reached, 65 = and at least one more)
                uint32_t hits_d, c;
                if (!UNSTAGED) {
                    const uint32_t col_s = (uint32_t)__cvta_generic_to_shared(col), wp_max = col_s + (uint32_t)(YASPH_MAXN + 1) * NB_ROW_BYTES;
                    // dynamic candidates, ascending slot == ascending sorted index (neighborhood_search.rs:353-366)
                    hits_d = (list_scan_candidates(cdyn_s, S.crun[0][lc], S.ncand[0][lc], q, a.g.radius_sq, col_s, wp_max) - col_s) / NB_ROW_BYTES;
                    const uint32_t cd0 = min(hits_d, (uint32_t)YASPH_MAXN);
                    // static candidates (neighborhood_search.rs:367-381)
                    c = (list_scan_candidates(cstat_s, S.crun[1][lc], S.ncand[1][lc], q, a.g.radius_sq, col_s + cd0 * NB_ROW_BYTES, wp_max) - col_s) / NB_ROW_BYTES;
                } else {
                    hits_d = list_scan_candidates_global<false>(a.pos, &tr, S.crun[0][lc], S.ncand[0][lc], q, a.g.radius_sq, col, 0u);
                    c = list_scan_candidates_global<true>(a.bpos, &tr, S.crun[1][lc], S.ncand[1][lc], q, a.g.radius_sq, col, min(hits_d, (uint32_t)YASPH_MAXN));
                }
                const uint32_t cd = min(hits_d, (uint32_t)YASPH_MAXN);
                const uint32_t ct = min(c, (uint32_t)YASPH_MAXN);
                // the reference's bookkeeping: "too many neighbors" when the 64th entry is written (:360,375); a static hit
                // with all 64 slots taken by dynamic neighbours indexes neighbor_set[64] and panics (:373) -- dropped here
                if (hits_d >= YASPH_MAXN) {
                    ++my_capped;
                    if (c > cd) ++my_dropped;
                } else if (c >= YASPH_MAXN) {
                    ++my_capped;
                }
                const uint32_t nkd = (cd + 3u) >> 2, nks = (ct - cd + 3u) >> 2;
                const unsigned long long own = h.own_lo + tl;
                const uint32_t* colw = &S.sl[tid][0];
                for (uint32_t kb = 0; kb < nkd; ++kb) {  // dynamic words, padded with the own slot
                    unsigned long long w = (unsigned long long)colw[2 * kb] | ((unsigned long long)colw[2 * kb + 1] << 32);
                    const uint32_t valid = cd - kb * 4u;  // >= 1
                    if (valid < 4u) {
                        const unsigned long long keep = (1ull << (16u * valid)) - 1ull;
                        w = (w & keep) | ((own * 0x0001000100010001ull) & ~keep);
                    }
                    a.lists[list_word_index(h.pstart, h.pcount, kb, tl)] = w;
                }
                for (uint32_t kb = 0; kb < nks; ++kb) {  // static words
                    unsigned long long w = 0ull;
#pragma unroll
                    for (uint32_t e = 0; e < 4; ++e)
                        if (cd + kb * 4 + e < ct) w |= (unsigned long long)col[cd + kb * 4 + e] << (16 * e);
                    a.lists[list_word_index(h.pstart, h.pcount, nkd + kb, tl)] = w;
                }
                a.counts[i] = cd | (ct << 8);
                my_nk = max(my_nk, nkd + nks);
                my_total += ct;
            }
        }
        LT_MARK(3)  // candidate scan + packing + stores
        my_nk = __reduce_max_sync(0xffffffffu, my_nk);
        if (lane_id() == 0 && my_nk) atomicMax(&S.nk_max, my_nk);
        __syncthreads();  // S.nk_max is complete; also orders this tile's reads of S.crun / S.ncand before the next tile's writes
        if (tid == 0 && mine) {
            a.tile_nk[t] = S.nk_max;
            cta_nk = max(cta_nk, S.nk_max);
        }
        pre.store(S.runs[(k + 2) % 3u], have2);
        LT_MARK(4)  // closing barrier
    }
#ifdef YASPH_LIST_TIMING
    if (lane_id() == 0 && a.dbg) {
        for (int i = 0; i < 5; ++i) atomicAdd(&a.dbg[i], (unsigned long long)tph[i]);
        atomicAdd(&a.dbg[5], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&a.dbg[6], (unsigned long long)k);
    }
#endif
    cp_async_wait_all();
    // statistics: warp reduce, one global atomic per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_total += __shfl_xor_sync(0xffffffffu, my_total, o);
        my_capped += __shfl_xor_sync(0xffffffffu, my_capped, o);
        my_dropped += __shfl_xor_sync(0xffffffffu, my_dropped, o);
    }
    if (lane_id() == 0) {
        S.wtotal[tid >> 5] = my_total;
        if (my_capped) atomicAdd(&a.ctl->capped, my_capped);
        if (my_dropped) atomicAdd(&a.ctl->dropped, my_dropped);
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long tot = 0;
        for (int w = 0; w < NB_THREADS / 32; ++w) tot += S.wtotal[w];
        if (tot) atomicAdd(&a.ctl->total_neighbors, tot);
        if (cta_nk > *(volatile unsigned int*)&a.ctl->max_nk) atomicMax(&a.ctl->max_nk, cta_nk);
    }
}

// Export to the reference's layout (neighborhood_search.rs:268-273,433-449): u16 counts + u32 global indices, stride 64.
__global__ void __launch_bounds__(256)
    k_export_lists(TileTables tt, const Control* __restrict__ ctl, const unsigned long long* __restrict__ lists, const uint32_t* __restrict__ counts,
                   uint16_t* __restrict__ out_cd, uint16_t* __restrict__ out_ct, uint32_t* __restrict__ out_lists) {
    __shared__ TileRuns tr;
    const uint32_t ntiles = ctl->num_tiles;
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        load_tile_runs(tr, tt.runs + t);
        __syncthreads();
        const TileHeader h = tr.hdr;
        for (uint32_t tl = threadIdx.x; tl < h.pcount; tl += blockDim.x) {
            const uint32_t i = h.pstart + tl;
            const uint32_t cw = counts[i];
            const uint2 c = make_uint2(cw & 0xFFu, (cw >> 8) & 0xFFu);
            out_cd[i] = (uint16_t)c.x;
            out_ct[i] = (uint16_t)c.y;
            if (out_lists) {
                const uint32_t nkd = (c.x + 3u) >> 2;
                for (uint32_t k = 0; k < c.y; ++k) {
                    const uint32_t e = k < c.x ? k : nkd * 4u + (k - c.x);  // static entries start at a word boundary
                    const uint32_t s = unpack_slot(lists[list_word_index(h.pstart, h.pcount, e >> 2, tl)], e);
                    out_lists[(size_t)i * YASPH_MAXN + k] = k < c.x ? dyn_slot_to_global(tr, s) : run_slot_to_global(tr.rs, s);
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace yasph
