/* frame_sched.h -- host-side scheduling rules of the render path (internal). */
#ifndef CLPT_FRAME_SCHED_H
#define CLPT_FRAME_SCHED_H

#ifdef __cplusplus
extern "C" {
#endif

int clpt_claim_direction(const unsigned long long *row_cost, int rows, int current, double *where);

#ifdef __cplusplus
}
#endif

#endif
