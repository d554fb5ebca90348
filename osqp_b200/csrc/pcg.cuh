// pcg.cuh -- state shared by the two device-resident PCG drivers:
//   pcg_graph.cu  lean 512-thread passes; the CG loop is a CUDA-graph WHILE node whose condition is
//                 set on the device (no host synchronisation); loop body = A pass, fused-operator
//                 pass with three dots, ONE x/r/p update.  Default from kGraphDriverMinNnz stored
//                 entries.  Also hosts the row-sharded (multi-GPU) driver, which runs the same
//                 kernels from a host loop with the exchanges in between.
//   pcg.cu        one persistent cooperative kernel per solve (grid.sync between phases); lowest
//                 launch count, default for small problems
#pragma once

#include "csr.cuh"

namespace b200 {

constexpr double kCgTolMin    = 1e-7;   // OSQP_CG_TOL_MIN    (osqp_api_constants.h:215)
constexpr double kCgPolishTol = 1e-5;   // OSQP_CG_POLISH_TOL (osqp_api_constants.h:216)
constexpr int    kAxResync    = 50;     // solves between exact recomputations of the carried A x
// stored entries (A + fused operator) from which the graph driver is the default: below, the
// ~8 launches of a graph solve cost more than the persistent kernel's lower occupancy
constexpr long long kGraphDriverMinNnz = 2000000;

enum { SLOT_RHS = 0, SLOT_RTY = 1, SLOT_RMAX = 2, SLOT_PKP = 3, SLOT_RKP = 4, SLOT_KPKP = 5, SLOT_COUNT = 6 };

struct PcgState {
  double    reduction_factor;
  double    eps_prev;
  double    last_eps;
  double    last_rnorm;
  long long total_iters;
  long long n_solves;
  int       zero_iters;
  int       last_iters;
};

struct PcgArgs {
  CsrView K2, A, At;
  int n, m;
  int n_shared;   // row-sharded, column-split layout: leading columns shared by several ranks (else 0)
  T *x, *p, *Kp, *r, *t, *b, *Ax, *w;
  const T* minv;
  const T* rho_vec;
  T rho;
  int admm_iter, max_iter, polishing, reduction_threshold, ax_valid;
  double prim_res, dual_res, tol_fraction;
  PcgState* st;
  double* red;   // SLOT_COUNT * gridDim.x
};


// running scalars of one solve in the graph driver (device memory)
struct PcgRun {
  double   eps, rTy, rnorm, pKp, beta, rhs_norm, alpha;
  double   dots[3];      // p'Kp, r'M^-1 Kp, Kp'M^-1 Kp (sharded driver: exchanged between the ranks)
  double   slots[16];    // column-split layout: (r'y, ||r||_inf) of every rank, gathered by ONE sum all-reduce
  double   rf, eps_prev;
  int      it, zero_iters;
  unsigned ticket[SLOT_COUNT];
};

}  // namespace b200

struct b200_pcg {
  const b200_csr* P  = nullptr;
  const b200_csr* A  = nullptr;
  const b200_csr* At = nullptr;
  b200_csr K2;
  int n = 0, m = 0;
  T *d_x = nullptr, *d_p = nullptr, *d_Kp = nullptr, *d_r = nullptr, *d_t = nullptr;
  T *d_Ax = nullptr, *d_w = nullptr;
  int ax_valid = 0, solves_since_sync = 0;
  T *d_minv = nullptr, *d_pd = nullptr, *d_ad = nullptr;
  const T* d_rho_vec = nullptr;
  T sigma = 0, rho = 0;
  int precond = 1, polishing = 0;
  b200::PcgState* d_state = nullptr;
  double*   d_red   = nullptr;
  int grid = 1, max_grid = 1;
  // graph driver / row-sharded driver
  int use_graph = 0;
  int sharded = 0;       // row-sharded multi-GPU mode (dist.cu)
  int lean = 0;          // loop body uses the flat 32-register passes
  int owner_slot = -1;   // library context (host thread) that created the solver; solves must come from it
  int p2p = 0;           // row-sharded: exchanges go through peer memory inside the kernels, loop is a graph
  int include_P = 1;     // sharded: only rank 0 carries P + sigma I in the fused operator
  b200::PcgArgs* d_args = nullptr;
  b200::PcgRun*  d_run  = nullptr;
  double*        d_gred = nullptr;     // SLOT_COUNT * gred_stride partials
  int            gred_stride = 0;
  void*          graph_exec = nullptr; // cudaGraphExec_t of the CG loop
  void*          graph      = nullptr;
};

// pcg_graph.cu
int  b200_pcg_graph_build(b200_pcg* s);
void b200_pcg_graph_destroy(b200_pcg* s);
int  b200_pcg_graph_solve(b200_pcg* s, const b200::PcgArgs& a);
int  b200_pcg_sharded_solve(b200_pcg* s, const b200::PcgArgs& a);
void b200_pcg_graph_configure_kernels();
void b200_pcg_profile_register(b200_pcg* s, bool alive);
