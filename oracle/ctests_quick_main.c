/* oracle/ctests_quick_main.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A short runner over the reference's own C test cases (S/test/problems/*.h, included from the reference tree in
 * place, nothing copied): the subset of S/test/run_tests.c that finishes in about a minute when every
 * scs_solve_lin_sys call crosses PCIe (level-2 plugin mode: the reference CORE compiled with -DINDIRECT=1 and the
 * five linsys.h functions taken from libscsb200.so, see oracle/Makefile ref_ctests_b200).  The full runner
 * (S/test/run_tests.c, 57 tests, ~9 minutes in this mode) is built next to it and run by the same GPU test when
 * SCS_B200_LONG_TESTS=1.  Prints the seconds every test took.
 */
#include <stdio.h>
#include <time.h>

#include "minunit.h"
#include "problem_utils.h"
#include "scs.h"

#include "problems/degenerate.h"
#include "problems/hs21_tiny_qp.h"
#include "problems/infeasible_lp.h"
#include "problems/infeasible_socp.h"
#include "problems/infeasible_tiny_qp.h"
#include "problems/lp_update.h"
#include "problems/qafiro_tiny_qp.h"
#include "problems/small_lp.h"
#include "problems/small_qp.h"
#include "problems/test_dual_exp_cone.h"
#include "problems/test_exp_cone.h"
#include "problems/test_mixed_cones.h"
#include "problems/test_power_cone.h"
#include "problems/test_root_plus.h"
#include "problems/test_soc_sizes.h"
#include "problems/test_box_cone.h"
#include "problems/test_psd_n1.h"
#include "problems/test_zero_cone.h"
#include "problems/unbounded_lp.h"
#include "problems/unbounded_socp.h"
#include "problems/unbounded_tiny_qp.h"
#include "problems/complex_PSD.h"
#include "problems/sd_and_complex_sd.h"
#include "problems/random_prob.h"
#include "problems/rob_gauss_cov_est.h"
#include "problems/hs21_tiny_qp_rw.h"
#include "problems/mpc_bug.h"

int tests_run = 0;

#define timed_test(test)                                                                  \
  do {                                                                                    \
    struct timespec t0_, t1_;                                                             \
    clock_gettime(CLOCK_MONOTONIC, &t0_);                                                 \
    mu_run_test(test);                                                                    \
    clock_gettime(CLOCK_MONOTONIC, &t1_);                                                 \
    scs_printf("[quick runner] %s: %.2f s\n", #test,                                      \
               (double)(t1_.tv_sec - t0_.tv_sec) + 1e-9 * (double)(t1_.tv_nsec - t0_.tv_nsec)); \
  } while (0)

static const char *all_tests(void) {
  timed_test(degenerate);
  timed_test(small_lp);
  timed_test(small_qp);
  timed_test(lp_update);
  timed_test(rob_gauss_cov_est);
  timed_test(complex_PSD);
  timed_test(sd_and_complex_sd);
  timed_test(hs21_tiny_qp);
  timed_test(hs21_tiny_qp_rw);
  timed_test(qafiro_tiny_qp);
  timed_test(infeasible_tiny_qp);
  timed_test(infeasible_lp);
  timed_test(infeasible_socp);
  timed_test(unbounded_tiny_qp);
  timed_test(unbounded_lp);
  timed_test(unbounded_socp);
  timed_test(random_prob);
  timed_test(mpc_bug);
  timed_test(test_exp_cone);
  timed_test(test_dual_exp_cone);
  timed_test(test_power_cone);
  timed_test(test_dual_power_cone);
  timed_test(test_soc_size1);
  timed_test(test_soc_size3);
  timed_test(test_multi_soc);
  timed_test(test_zero_cone);
  timed_test(test_box_cone_lp);
  timed_test(test_psd_n1);
  timed_test(test_mixed_cones);
  timed_test(test_root_plus_equivalence);
  return 0;
}

int main(void) {
  const char *result = all_tests();
  if (result != 0) {
    scs_printf("%s\n", result);
    scs_printf("TEST FAILED!\n");
  } else {
    scs_printf("ALL TESTS PASSED\n");
  }
  scs_printf("Tests run: %d\n", tests_run);
  return result != 0;
}
