// solver_internal.cuh -- host helpers of scs_solver.cu that the batch engine (batch.cu) shares.
#pragma once
#include "common.cuh"

namespace b200 {

// validate_cones + validate_lin_sys + validate settings (scs.c:364-429, scs_matrix.c:65-131,
// cones.c:583-755); prints the reference's message and returns < 0 on the first violation.
int validate_problem(const ScsData *d, const ScsCone *k, const ScsSettings *stgs);
// populate_on_failure, scs.c:316-345: NaN-fill sol / info, allocate sol vectors when absent.
void populate_on_failure(int m, int n, ScsSolution *sol, ScsInfo *info, int status_val, const char *msg);

}  // namespace b200
