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

/* Claim ORDER: the rows of the costliest band first, then the rest.
 *
 * The direction rule above makes a frame END on its cheap side.  On short frames that is not
 * enough: the rows under the horizon hold rays that graze the whole terrain (thousands of
 * triangle tests each on a shallow tree), one such tile runs for a large part of the frame,
 * and when it is claimed half-way through, the frame ends with a handful of warps finishing
 * their tiles on an otherwise idle GPU.  Claimed FIRST, the same tiles run beside all the
 * cheap ones.  "Hot" rows are those whose cost (3-row mean) exceeds `hot_factor` times the
 * mean row cost, at most a quarter of the rows; they keep their screen order (neighbouring
 * claims stay neighbours in the tree -- sorting all rows by cost was measured in round 1 and
 * loses more in locality than it wins), and so do the others, which follow in the given
 * direction.
 *
 *   order     out: order[k] = screen row claimed k-th (a permutation of 0..rows-1)
 *   reverse   direction of the rows that are not hot (clpt_claim_direction)
 * Returns the number of hot rows (0: plain order in the given direction).
 */
int
clpt_claim_order(const unsigned long long *row_cost, int rows, int reverse, double hot_factor, int *order) {
    int n_hot = 0;
    if (rows <= 0) {
        return 0;
    }
    unsigned char hot[4096];
    const int can_mark = rows >= 16 && rows <= (int)sizeof(hot);
    if (can_mark) {
        double total = 0.0;
        for (int i = 0; i < rows; i++) {
            total += (double)row_cost[i];
        }
        const double mean = total / rows;
        int budget = rows / 4;
        for (int i = 0; i < rows; i++) {
            const double a = (double)row_cost[i > 0 ? i - 1 : i], b = (double)row_cost[i],
                         c = (double)row_cost[i + 1 < rows ? i + 1 : i];
            hot[i] = total > 0.0 && (a + b + c) / 3.0 > hot_factor * mean;
            n_hot += hot[i];
        }
        while (n_hot > budget) { /* keep the costliest: drop the cheapest marked row */
            int drop = -1;
            for (int i = 0; i < rows; i++) {
                if (hot[i] && (drop < 0 || row_cost[i] < row_cost[drop])) {
                    drop = i;
                }
            }
            hot[drop] = 0;
            n_hot--;
        }
    }
    int k = 0;
    if (can_mark) {
        for (int i = 0; i < rows; i++) {
            if (hot[i]) {
                order[k++] = i;
            }
        }
    }
    for (int j = 0; j < rows; j++) {
        const int i = reverse ? rows - 1 - j : j;
        if (!can_mark || !hot[i]) {
            order[k++] = i;
        }
    }
    return n_hot;
}
