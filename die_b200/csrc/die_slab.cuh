// die_slab.cuh -- one field split into row slabs over G GPUs (mode 2).
//
// Rank r owns rows [r*rows_per, (r+1)*rows_per) of every per-cell array (medium A/B, claim table,
// consumed_field, gradient cache) and a set of agent slots (two global ranges: its slab's alive
// agents, which the reference creates in row-major cell order, and an even share of the ghost
// slots).  Every rank allocates its slabs as symmetric memory, so each kernel sees a table of G
// peer base pointers and reads / atomically updates REMOTE cells directly over NVLink (P2P
// ld / st / red): the halo rows of the stencil, the sensed gradient of an agent that looks across a
// slab seam, the claim of an agent that walked into the neighbour's slab.  There is no packing, no
// staging buffer and no agent migration; ranks only meet at three barriers per step.
//
// With G = 1 these accessors are not used at all (the kernels are instantiated with SLAB = false
// and plain pointers).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace die {

constexpr int kMaxRanks = 8;

struct SlabGeom {
    int G, rank;
    int H, W;                  // GLOBAL field
    int rows_per;              // H / G
    int slab_shift;            // log2(slab_cells) when it is a power of two, else -1
    int slab_cells;            // rows_per * W  (< 2^31)
    int64_t M;                 // GLOBAL slot count
    // slot ownership: rank q owns global slots [s0[q], s0[q]+n0[q]) (local 0..n0-1) and then
    // [s1[q], s1[q]+n1[q]) (local n0..n0+n1-1)
    int64_t s0[kMaxRanks], n0[kMaxRanks], s1[kMaxRanks], n1[kMaxRanks];
};

struct SlabTables {            // device arrays [G] of peer base pointers (symmetric memory)
    double* const* medium_in;  // each [3][rows_per][W]
    double* const* medium_out;
    int32_t* const* claim;     // each [rows_per*W]
    double* const* consumed;
    double2* const* grad;
    double* const* action;     // each [3][Ml(q)]
    // Corner mirror (LOCAL memory, may be null).  The reference creates ~90 % of its slots as "ghosts" at
    // (0, 0) and positions wrap, so for thousands of steps they live within a few hundred cells of the four
    // corners of the field -- on ranks 0 and G-1 -- and every other rank's ghost gathers would cross NVLink
    // into those two.  The corner_r x corner_r patches at the four corners of the published gradient, the
    // current env_food and consumed_field are therefore copied to every rank after each field pass.
    // Layout [2 corner_r][2 corner_r]: patch row = row (top) or row - (H - 2 corner_r) (bottom), same for columns.
    const double2* corner_grad;
    const double* corner_food;   // env_food of the CURRENT medium (medium_in of the next forward)
    const double* corner_cons;
    int corner_r;
};

// index into the corner mirror of global cell `cell`, or -1
__device__ __forceinline__ int slab_corner_index(const SlabTables& t, int H, int W, int cell) {
    const int R = t.corner_r;
    const int row = cell / W, col = cell - row * W;
    const bool top = row < R, left = col < R;
    if ((top || row >= H - R) && (left || col >= W - R))
        return (top ? row : row - (H - 2 * R)) * (2 * R) + (left ? col : col - (W - 2 * R));
    return -1;
}

__device__ __forceinline__ int slab_owner(const SlabGeom& g, int cell, int& local) {
    const int o = (g.slab_shift >= 0) ? (cell >> g.slab_shift) : (cell / g.slab_cells);
    local = cell - o * g.slab_cells;
    return o;
}

// element `cell` (global linear index) of channel `ch` of a [3][rows_per][W] slab set
__device__ __forceinline__ double* slab_chan(double* const* tab, const SlabGeom& g, int ch, int cell) {
    int local;
    const int o = slab_owner(g, cell, local);
    return (double*)__ldg((const unsigned long long*)(tab + o)) + (int64_t)ch * g.slab_cells + local;
}

template <typename T>
__device__ __forceinline__ T* slab_cell(T* const* tab, const SlabGeom& g, int cell) {
    int local;
    const int o = slab_owner(g, cell, local);
    return (T*)__ldg((const unsigned long long*)(tab + o)) + local;
}

// read-only gathers of the forward / feed kernels: the local corner mirror when the cell is in it
__device__ __forceinline__ double2 slab_load_grad(const SlabTables& t, const SlabGeom& g, int cell) {
    if (t.corner_grad != nullptr) {
        const int b = slab_corner_index(t, g.H, g.W, cell);
        if (b >= 0) return __ldg(t.corner_grad + b);
    }
    return __ldg(slab_cell(t.grad, g, cell));
}

__device__ __forceinline__ double slab_load_food(const SlabTables& t, const SlabGeom& g, int cell) {
    if (t.corner_food != nullptr) {
        const int b = slab_corner_index(t, g.H, g.W, cell);
        if (b >= 0) return __ldg(t.corner_food + b);
    }
    return __ldg(slab_chan(t.medium_in, g, 1, cell));
}

__device__ __forceinline__ double slab_load_consumed(const SlabTables& t, const SlabGeom& g, int cell) {
    if (t.corner_cons != nullptr) {
        const int b = slab_corner_index(t, g.H, g.W, cell);
        if (b >= 0) return __ldg(t.corner_cons + b);
    }
    return __ldg(slab_cell(t.consumed, g, cell));
}

__device__ __forceinline__ int64_t slab_slots_of(const SlabGeom& g, int q) { return g.n0[q] + g.n1[q]; }

__device__ __forceinline__ int64_t slab_slot_global(const SlabGeom& g, int q, int64_t local) {
    return (local < g.n0[q]) ? g.s0[q] + local : g.s1[q] + (local - g.n0[q]);
}

// owner rank and local index of a global slot id
__device__ __forceinline__ int slab_slot_owner(const SlabGeom& g, int64_t gid, int64_t& local) {
    {   // almost always the winner of a cell of this slab is one of this rank's own alive slots
        const int r = g.rank;
        if (gid >= g.s0[r] && gid < g.s0[r] + g.n0[r]) {
            local = gid - g.s0[r];
            return r;
        }
    }
    for (int q = 0; q < g.G; ++q) {
        if (gid >= g.s0[q] && gid < g.s0[q] + g.n0[q]) {
            local = gid - g.s0[q];
            return q;
        }
        if (gid >= g.s1[q] && gid < g.s1[q] + g.n1[q]) {
            local = g.n0[q] + (gid - g.s1[q]);
            return q;
        }
    }
    local = 0;
    return 0;
}

}  // namespace die
