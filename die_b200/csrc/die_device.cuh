// die_device.cuh -- device-side arithmetic shared by the agent and field kernels.
//
// Every function here restates one numpy / pandas expression of the reference in IEEE
// float64 with the SAME operation order, so integer results (cell indices, turn
// decisions) are bit-exact.  The library is compiled with -fmad=false: no mul+add is ever
// contracted into an FMA, as in numpy.  Citations are file:line under /root/reference.
#pragma once
#include <cstdint>
#include <cfloat>
#include <cuda_runtime.h>

#include "die_math.h"
#include "die_turn.h"

namespace die {

constexpr double kPi    = DIE_PI;        // np.pi
constexpr double kTwoPi = DIE_TWO_PI;    // 2 * np.pi (exact doubling)

// ---- nearest grid cell ------------------------------------------------------------------
// xarray .sel(method='nearest') -> pandas.Index.get_indexer(method='nearest') on
// np.linspace(0, 1, n)  (core/utils.py:39-54, core/data_init.py:99-100).
// g(i) = i * (1/(n-1)), g(n-1) = 1.0 exactly.  L = last g <= c, R = first g >= c; L wins iff
// (c - g[L]) < (g[R] - c) strictly; below/above range clamps.
struct Axis {
    int    n;
    double step;     // 1.0 / (n - 1)
    double nm1;      // (double)(n - 1)
};

inline Axis make_axis(int n) {                 // host side: passed to the kernels by value
    Axis a;
    a.n = n;
    a.nm1 = (double)(n - 1);
    a.step = 1.0 / a.nm1;
    return a;
}

__device__ __forceinline__ double grid_coord(const Axis& a, int i) {
    return (i >= a.n - 1) ? 1.0 : __dmul_rn((double)i, a.step);
}

// Straight-line form, no conversions and no data-dependent branch on the hot path.
//   i  = rint(c (n-1) - 0.5) by the add-magic trick: the integer is the low word of the sum, its
//        double the sum minus the magic.  i estimates L and can be off by one -- but only when c
//        lies within a few ulp of a grid coordinate, and then that coordinate IS the nearest
//        one and the comparison below still lands on it (checked exhaustively against pandas in
//        tests/test_gpu_cells.py), so no correction step is needed.
//   gl = g(i), gr = g(i+1) with the linspace quirk g(n-1) = 1.0 exactly.
//   out-of-range c gives i < 0 or i > n-2; the final clamp resolves those to 0 / n-1.
// gl == c needs no special case (then 0 < gr - c picks i).
__device__ __forceinline__ int nearest_cell(double c, const Axis& a) {
    if (!(fabs(c) < 4.0)) return (c >= 4.0) ? a.n - 1 : 0;            // absurd / NaN: clamp (never hot)
    const double u = __dadd_rn(__dsub_rn(__dmul_rn(c, a.nm1), 0.5), DIE_RINT_MAGIC);
    const int i = __double2loint(u);
    const double di = __dsub_rn(u, DIE_RINT_MAGIC);
    const double gl = __dmul_rn(di, a.step);
    const double gr = (i == a.n - 2) ? 1.0 : __dmul_rn(__dadd_rn(di, 1.0), a.step);
    const int r = (__dsub_rn(c, gl) < __dsub_rn(gr, c)) ? i : i + 1;
    return min(max(r, 0), a.n - 1);
}

// The same cell from a coordinate known only to +-guard cells (the sense position formed with the float32 cosine of
// die_sincosf_approx): frac = distance of c (n-1) above the integer below it; the nearest grid coordinate changes at
// frac = 0.5, so the answer is robust -- equal to nearest_cell of the exact coordinate -- iff |frac - 0.5| > guard
// (nearest_cell's own comparison errs by ~1e-13 cells: `guard` includes 1e-9 for it).  Returns false where the caller
// must evaluate the exact coordinate instead (also for NaN / absurd coordinates).
__device__ __forceinline__ bool nearest_cell_guarded(double c, const Axis& a, double guard, int* cell) {
    const double t = __dmul_rn(c, a.nm1);
    const double u = __dadd_rn(__dsub_rn(t, 0.5), DIE_RINT_MAGIC);
    const int i = __double2loint(u);
    const double frac = __dsub_rn(t, __dsub_rn(u, DIE_RINT_MAGIC));          // in [0, 1]
    *cell = min(max(frac < 0.5 ? i : i + 1, 0), a.n - 1);
    return fabs(c) < 4.0 && fabs(__dsub_rn(frac, 0.5)) > guard;
}

// ---- numpy float remainder, renormalize_radians, np.angle, nan_to_num ------------------------
// The turn-rule arithmetic lives in die_turn.h (host + device, so its guard-band logic is testable on
// the CPU); these are the device-side names the kernels use.
__device__ __forceinline__ double np_remainder(double a, double b) { return die_np_remainder(a, b); }

// coords % 1.  (core/env.py:155) -- returns exactly 1.0 for tiny negatives.
__device__ __forceinline__ double mod1(double a) { return die_np_remainder(a, 1.0); }

// renormalize_radians (core/utils.py:177-179): (r - pi) % (-2 pi) + pi, in (-pi, pi].
__device__ __forceinline__ double renormalize_radians(double r) { return die_renormalize_radians(r); }

// np.angle(x + np.multiply(1j, y))  (core/utils.py:158-168); FAST = die_atan2_fast (<= 1.8 ulp) for
// angles that are only compared against thresholds, else the compensated die_atan2.
template <bool FAST>
__device__ __forceinline__ double angle_xy(double x, double y) { return die_angle_xy(x, y, FAST ? 1 : 0); }

// ---- Philox4x32-10 counter RNG (perf mode; validation mode injects the host's draws) -------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

__device__ __forceinline__ uint4 philox_draw(uint64_t seed, uint64_t step, uint64_t slot, uint32_t stream) {
    // counter = (slot lo, slot hi, step lo, step hi ^ stream<<24); key = seed
    const uint4 ctr = make_uint4((uint32_t)slot, (uint32_t)(slot >> 32), (uint32_t)step,
                                 (uint32_t)(step >> 32) ^ (stream << 24));
    return philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// uniform in [0, 1) with 53 random bits, like numpy's random_sample
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (double)(v >> 11) * (1.0 / 9007199254740992.0);
}

// uniform in [0, 1) with 32 random bits
__device__ __forceinline__ double u32(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

// ---- block reduction (fixed order => deterministic) ------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace die
