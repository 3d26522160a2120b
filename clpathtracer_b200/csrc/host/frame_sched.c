/* frame_sched.c -- host-side scheduling rules of the render path (no device code).
 *
 * Kept in plain C next to the other host logic so that the CPU test suite can
 * drive them through the library (tests/test_host_parity.py) without a GPU.
 */
#include "frame_sched.h"

/* Claim direction of the persistent render kernel (csrc/cuda/render_kernel.cu).
 *
 * A warp tile is a whole pixel's samples and bounces, so what matters at the end
 * of a frame is the longest tile still running; cheap rows claimed after the
 * expensive ones give no cover (sky tiles take microseconds, the grazing rows
 * under the horizon up to a millisecond).  Given the per-row cost of the
 * previous frame, find the costliest band (5-row moving sum) and start from the
 * end it is nearer to.  The band has to sit clearly in the far half (beyond 55%
 * of the way) to flip the direction, so noise does not.
 *
 *   row_cost  cost of each row of blocks, in SCREEN order (row 0 first)
 *   rows      number of rows; fewer than 16 keeps `current`
 *   current   0 = claims run top-down (row 0 first), 1 = bottom-up
 *   where     optional out: position of the costliest band, 0..1 in screen order
 * Returns the direction for the next frame.
 */
int
clpt_claim_direction(const unsigned long long *row_cost, int rows, int current, double *where) {
    if (where) {
        *where = 0.0;
    }
    if (rows < 16) {
        return current;
    }
    unsigned long long best = 0, sum = 0;
    int best_at = 0;
    for (int i = 0; i < rows; i++) {
        sum += row_cost[i];
        if (i >= 5) {
            sum -= row_cost[i - 5];
        }
        if (sum > best) {
            best = sum;
            best_at = i - 2;
        }
    }
    if (best == 0) {
        return current; /* nothing measured */
    }
    const double at = (double)best_at / (double)(rows - 1);
    if (where) {
        *where = at;
    }
    if (!current && at > 0.55) {
        return 1;
    }
    if (current && at < 0.45) {
        return 0;
    }
    return current;
}
