/* turn_check.c -- host build of die_b200/csrc/die_turn.h for tests/test_turn_quick.py:
 * evaluates the guard-banded quick turn decision and the exact (reference-arithmetic) one on
 * arrays, so the test can require "decided => identical" on adversarial inputs without a GPU.
 * Build: gcc -O2 -std=c99 -ffp-contract=off -shared -fPIC. */
#include "../../die_b200/csrc/die_turn.h"

/* out[i*5 + {0: decided, 1: quick turn, 2: quick mask, 3: exact turn, 4: exact mask}] */
int die_turn_check(long n, const double* gx, const double* gy, const double* theta,
                   int normalized, int use_clip, double clip, double atol, double sense_radians,
                   int* out) {
    const die_turn_plan_t plan = die_turn_plan(normalized, use_clip, clip, atol, sense_radians);
    for (long i = 0; i < n; ++i) {
        double sn, cs;
        die_sincos(theta[i], &sn, &cs);
        die_turn_t q;
        q.turn = 99;
        q.deposit_mask = 99;
        const int decided = plan.enabled && die_turn_quick(&plan, gx[i], gy[i], sn, cs, theta[i], atol, sense_radians, &q);
        double ngx = gx[i], ngy = gy[i];
        die_normalize_gradient(&ngx, &ngy, normalized, use_clip, clip);
        const die_turn_t e = die_turn_exact(ngx, ngy, theta[i], atol, sense_radians);
        out[i * 5 + 0] = decided;
        out[i * 5 + 1] = q.turn;
        out[i * 5 + 2] = q.deposit_mask;
        out[i * 5 + 3] = e.turn;
        out[i * 5 + 4] = e.deposit_mask;
    }
    return plan.enabled;
}

/* die_sqrt_near (die_math.h) on an array: y_out[i] = its result, fast_out[i] = 1 where the Newton step was accepted.
 * Returns whether the plan for r0 is enabled. */
int die_sqrt_near_check(long n, const double* s, double r0, double* y_out, int* fast_out) {
    const die_sqrt_near_t plan = die_sqrt_near_plan(r0);
    for (long i = 0; i < n; ++i) {
        y_out[i] = die_sqrt_near(&plan, s[i]);
        int fast = 0;
        if (plan.enabled) {
            const double e = fma(-plan.r0, plan.r0, s[i]);
            const double y = fma(e, plan.half_inv, plan.r0);
            const double d = fma(-y, y, s[i]);
            fast = fabs(e) < plan.elim && fabs(d) < plan.margin;
        }
        fast_out[i] = fast;
    }
    return plan.enabled;
}
