/* oracle/ctests_quick_main.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A table-driven runner over the reference's own C test cases (S/test/problems/<case>.h, included from the reference
 * tree in place, nothing copied): the cases of S/test/run_tests.c that finish in about a minute when every
 * scs_solve_lin_sys call crosses PCIe (level-2 plugin mode: the reference CORE compiled with -DINDIRECT=1 and the
 * five linsys.h functions taken from libscsb200.so, see oracle/Makefile ref_ctests_b200).  The reference's full
 * runner (S/test/run_tests.c, 57 tests, ~9 minutes in this mode) is built next to it and run by the same GPU test
 * when SCS_B200_LONG_TESTS=1.
 *
 *   run_tests_b200_linsys_quick [substring ...]   run the cases whose name contains one of the substrings (all if none)
 *
 * Output: one "[quick runner] <case>: <seconds> s  ok|FAILED" line per case, a table of the slowest cases, then the
 * verdict lines the reference's runner prints ("ALL TESTS PASSED" / "TEST FAILED!", "Tests run: N").
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "minunit.h"
#include "problem_utils.h"
#include "scs.h"

/* tiny and small problems with known answers */
#include "problems/small_lp.h"
#include "problems/small_qp.h"
#include "problems/lp_update.h"
#include "problems/degenerate.h"
#include "problems/hs21_tiny_qp.h"
#include "problems/hs21_tiny_qp_rw.h"
#include "problems/qafiro_tiny_qp.h"
/* certificates */
#include "problems/infeasible_tiny_qp.h"
#include "problems/infeasible_lp.h"
#include "problems/infeasible_socp.h"
#include "problems/unbounded_tiny_qp.h"
#include "problems/unbounded_lp.h"
#include "problems/unbounded_socp.h"
/* one cone type at a time */
#include "problems/test_zero_cone.h"
#include "problems/test_box_cone.h"
#include "problems/test_soc_sizes.h"
#include "problems/test_psd_n1.h"
#include "problems/test_exp_cone.h"
#include "problems/test_dual_exp_cone.h"
#include "problems/test_power_cone.h"
#include "problems/test_mixed_cones.h"
#include "problems/test_root_plus.h"
/* semidefinite programs (LAPACK in the reference core) and data-file problems */
#include "problems/rob_gauss_cov_est.h"
#include "problems/complex_PSD.h"
#include "problems/sd_and_complex_sd.h"
#include "problems/random_prob.h"
#include "problems/mpc_bug.h"

int tests_run = 0; /* referenced by minunit.h users inside the case headers */

typedef const char *(*case_fn)(void);
struct quick_case {
  const char *name;
  case_fn run;
  double seconds;
  int failed;
};

#define CASE(fn) {#fn, fn, 0.0, 0}
static struct quick_case cases[] = {
    CASE(small_lp),          CASE(small_qp),           CASE(lp_update),         CASE(degenerate),
    CASE(hs21_tiny_qp),      CASE(hs21_tiny_qp_rw),    CASE(qafiro_tiny_qp),    CASE(infeasible_tiny_qp),
    CASE(infeasible_lp),     CASE(infeasible_socp),    CASE(unbounded_tiny_qp), CASE(unbounded_lp),
    CASE(unbounded_socp),    CASE(test_zero_cone),     CASE(test_box_cone_lp),  CASE(test_soc_size1),
    CASE(test_soc_size3),    CASE(test_multi_soc),     CASE(test_psd_n1),       CASE(test_exp_cone),
    CASE(test_dual_exp_cone), CASE(test_power_cone),   CASE(test_dual_power_cone), CASE(test_mixed_cones),
    CASE(test_root_plus_equivalence), CASE(rob_gauss_cov_est), CASE(complex_PSD), CASE(sd_and_complex_sd),
    CASE(random_prob),       CASE(mpc_bug),
};
static const int ncases = (int)(sizeof(cases) / sizeof(cases[0]));

static double now_s(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static int selected(const char *name, int argc, char **argv) {
  int a;
  if (argc <= 1) return 1;
  for (a = 1; a < argc; ++a)
    if (strstr(name, argv[a])) return 1;
  return 0;
}

static int by_time_desc(const void *x, const void *y) {
  const double a = (*(const struct quick_case *const *)x)->seconds, b = (*(const struct quick_case *const *)y)->seconds;
  return a < b ? 1 : (a > b ? -1 : 0);
}

int main(int argc, char **argv) {
  const struct quick_case *order[sizeof(cases) / sizeof(cases[0])];
  const char *first_message = 0;
  int i, nrun = 0, nfailed = 0;
  double total = 0.0;
  for (i = 0; i < ncases; ++i) {
    struct quick_case *c = &cases[i];
    const char *message;
    double t0;
    if (!selected(c->name, argc, argv)) continue;
    t0 = now_s();
    message = c->run();
    c->seconds = now_s() - t0;
    c->failed = message != 0;
    total += c->seconds;
    order[nrun++] = c;
    scs_printf("[quick runner] %s: %.2f s  %s\n", c->name, c->seconds, c->failed ? "FAILED" : "ok");
    if (c->failed) {
      ++nfailed;
      scs_printf("[quick runner]   %s\n", message);
      if (!first_message) first_message = message;
    }
  }
  qsort(order, (size_t)nrun, sizeof(order[0]), by_time_desc);
  scs_printf("[quick runner] %d cases in %.1f s; slowest:", nrun, total);
  for (i = 0; i < nrun && i < 5; ++i) scs_printf(" %s %.1f s%s", order[i]->name, order[i]->seconds, i + 1 < nrun && i < 4 ? "," : "");
  scs_printf("\n");
  if (nfailed) {
    scs_printf("%s\n", first_message);
    scs_printf("TEST FAILED!\n");
  } else {
    scs_printf("ALL TESTS PASSED\n");
  }
  scs_printf("Tests run: %d\n", nrun);
  return nfailed != 0;
}
