// Device-side ActionIdentify (declarations).  See action.cu.
#pragma once
#include "common.cuh"

namespace ydst {

struct ActionIdentifyDev;
// kinds: 0 TakeOff, 1 Landing, 2 Glide (p0, p1 = delta x, y), 3 FastCrossing (p0 = speed), 4 BreakInto (p0 = timeout)
ActionIdentifyDev* action_create(int max_age, int max_size, const int* kinds, const int* class_ids, const double* p0, const double* p1, int n_rules, int cap);
void action_destroy(ActionIdentifyDev* a);
// rows_host (K,6) int32; triples_out (track id, class id, rule index) in the reference's order; returns their number.  Synchronises.
int action_update(ActionIdentifyDev* a, const int32_t* rows_host, int K, double now, int32_t* triples_out, cudaStream_t stream);

}  // namespace ydst
