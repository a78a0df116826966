// gslam_b200/csrc/ba_internal.cuh — host-side types and entry points shared by the bundle-adjustment translation units
// (ba.cu: graph upload, kernels of the sweep / Schur / local PCG paths; ba_pcg_bcsr.cu: the multi-CTA block-CSR PCG of large
// reduced systems; ba_dist.cu: the landmark-sharded multi-GPU solve and its communicator).  Not part of the C-ABI.
#pragma once
#include <vector>

#include "ba_device.cuh"
#include "common.cuh"

struct gb_ba_graph {
  BaDev d{};
  uint8_t* slab = nullptr;  // one device allocation (or the ctx arena) holding everything below
  size_t slab_bytes = 0;
  bool from_arena = false;
  double *pose_init = nullptr, *pts_init = nullptr, *pose_wc_out = nullptr;
  double* buf = nullptr;     // internal [S | gt | diagU | cost | pad]
  double* d_cost = nullptr;  // internal candidate cost
  size_t buf_doubles = 0;
  gb_ba_options opt{};
  std::vector<int> sorted_to_orig;  // sorted observation slot -> caller's edge index
  std::vector<int> cam_perm_h;      // host copy of cam_perm (camera-sorted slot -> sorted slot): refresh of a cached graph
  double* pose_wc_in = nullptr;     // device: the uploaded T_wc poses (blob), source of ba_prepare_kernel
  bool begun = false;
  // PCG dispatch: single-CTA block-sparse kernel, else one-cluster kernel (pcg_cluster = 8/16), else generic multi-kernel
  bool pcg_sparse = false;
  size_t pcg_sparse_smem = 0;
  int pcg_max_row_blocks = 0;  // longest block row of S
  int sweep_mode = 0;           // 0 = by size (ba_sweep.cu from 64k observations), 1 = ba.cu's latency-tuned kernel, 2 = ba_sweep.cu
  bool sweep_only = false;     // the last begin came from gb_ba_graph_sweep and nothing else ran since
  int pcg_nact = 0;            // cameras with at least one free dof (the sparse PCG kernel gives lanes to these only)
  int pcg_cluster = 0;
  size_t pcg_smem = 0;
  // multi-CTA block-CSR PCG (ba_pcg_bcsr.cu): plan made at graph creation when the block structure exists and the single-CTA
  // kernel does not apply
  bool pcg_bcsr = false;
  int bcsr_ctas = 0, bcsr_K = 0, bcsr_in_smem = 0, bcsr_max_cams = 0, bcsr_max_blocks = 0, bcsr_cluster = 0, bcsr_blk_stride = 37;
  size_t bcsr_smem = 0;
  unsigned short bcsr_need[16] = {0};  // cluster mode: bcsr_need[c] = CTAs that read the rows of u owned by CTA c
  int* bcsr_cta_cam = nullptr;   // device [bcsr_ctas + 1]: first camera of each CTA's block-row range (base of one allocation)
  double* bcsr_part = nullptr;   // device [bcsr_ctas * 2]: per-CTA (gamma, delta) partials
  double* bcsr_u = nullptr;      // device [n6]: the published u = Minv r
  unsigned int* bcsr_bar = nullptr;  // device: grid barrier counter
  double* rbuf = nullptr;        // device: compact reduced system [Sb (nnzb*36) | g~ | diag U | cost | pad]
  size_t rbuf_doubles = 0;
  // direct solver (ba_chol.cu): skyline plan [first | rowoff | last] uploaded with the blob, when the skyline fits one SM
  bool chol_ok = false;
  const int* chol_plan = nullptr;
  int chol_blocks = 0;
  size_t chol_smem = 0;
  void* sp_alloc = nullptr;      // device allocation holding the landmark-chunk Schur plan + staging (BaDev::sp_*), or null
  void* pe_alloc = nullptr;      // device allocation holding the pose-graph edges and their gather plans (BaDev::pe_* / pc_* / pp_*)
  void* sw_alloc = nullptr;      // device allocation holding the large-graph sweep's item plan (BaDev::sw_*), made on first use
  std::vector<int> pt_off_h, cam_off_h;  // host copies of pt_off / cam_off (the sweep plan is cut from them)
  // landmark shard (multi-GPU global BA): this graph holds landmarks [shard_lo, shard_hi) of the caller's problem
  int shard_lo = 0, shard_hi = 0, shard_rank = 0, shard_world = 1;
};

// ---- ba.cu ------------------------------------------------------------------------------------------------------------------
// shard_world > 1: keep only the landmarks of `shard_rank` (contiguous range balanced by observation count) and their edges;
// all cameras and the GLOBAL covisibility block structure are kept, so that every rank's reduced system has the same layout.
int ba_graph_create_impl(gb_ctx* ctx, const gb_ba_problem* pb, gb_ba_graph** out, bool use_arena, int shard_rank, int shard_world,
                         const gb_pose_edges* pose_edges = nullptr);
// one LM iteration in pieces, on the COMPACT reduced layout rbuf = [Sb | g~ | diag U | cost | pad] (g->rbuf_doubles doubles):
int ba_reduce_local_compact(gb_ctx* ctx, gb_ba_graph* g, double* rbuf);           // sweep (if needed) + Schur blocks of the shard
int ba_backsub_cost_compact(gb_ctx* ctx, gb_ba_graph* g, double* d_cost);         // back-substitution + candidate cost of the shard
int ba_commit_compact(gb_ctx* ctx, gb_ba_graph* g, const double* rbuf, const double* d_cost);  // LM accept / reject + install
int ba_read_result(gb_ctx* ctx, gb_ba_graph* g, gb_ba_result* res);               // one sync; fills res from the device scalars

// ---- ba_pose.cu -------------------------------------------------------------------------------------------------------------
// pose-graph terms (SE3Edge / GPSEdge, Optimizer.h:127-148) on the stepwise dense-layout solver path
int ba_pose_validate(gb_ctx* ctx, const gb_ba_problem* pb, const gb_pose_edges* pe);
int ba_pose_attach(gb_ctx* ctx, gb_ba_graph* g, const gb_pose_edges* pe);   // upload edges + gather plans (graph creation)
void ba_pose_free(gb_ba_graph* g);
int ba_pose_linearize(gb_ctx* ctx, gb_ba_graph* g, cudaStream_t s);         // after the sweep: records -> U, g_c, cost terms
int ba_pose_offdiag(gb_ctx* ctx, gb_ba_graph* g, double* buf, cudaStream_t s);  // after the Schur complement: S_ij += J_i' Omega J_j
int ba_pose_cost(gb_ctx* ctx, gb_ba_graph* g, cudaStream_t s);              // candidate cost terms at pose_new

// ---- ba_sweep.cu ------------------------------------------------------------------------------------------------------------
// the bandwidth-tuned residual + Jacobian sweep of large graphs (persistent CTAs, pose table in shared memory, bulk-copied W tiles)
void ba_sweep_plan_host(int nc, int np, int cam_split, const std::vector<int>& cam_off, const std::vector<int>& pt_off, int n_teams,
                        std::vector<int>& items4, std::vector<int>& team_off);  // host-only: item records + per-team ranges
void ba_sweep_plan_drop(gb_ba_graph* g);  // (cam_split changed / graph destroyed)
int ba_sweep_launch(gb_ctx* ctx, gb_ba_graph* g, const BaDev& d, cudaStream_t s, int which /* 3 whole, 1 cameras, 2 landmarks */);

// ---- ba_pcg_bcsr.cu ---------------------------------------------------------------------------------------------------------
// Decide whether / how the multi-CTA block-CSR PCG applies to `g` (fills the bcsr_* fields, allocates its small device buffers).
int ba_pcg_bcsr_plan(gb_ctx* ctx, gb_ba_graph* g, const int* s_rowptr_host, const int* s_col_host);
void ba_pcg_bcsr_free(gb_ba_graph* g);
// damp + block-Jacobi PCG on the reduced camera system held in `rbuf` + retraction of the cameras (pose_new, Rt_new, x)
int ba_pcg_bcsr_launch(gb_ctx* ctx, gb_ba_graph* g, const double* rbuf);

// ---- ba_chol.cu -------------------------------------------------------------------------------------------------------------
bool ba_chol_plan_host(gb_ctx* ctx, int nc, const int* s_rowptr, const int* s_col, std::vector<int>& plan3, int* nblocks, size_t* smem);
// block-skyline Cholesky solve of the damped reduced camera system in the dense-layout `buf` (+ g->d.Sb block values) + retraction
int ba_chol_launch(gb_ctx* ctx, gb_ba_graph* g, double* buf, bool pdl);

// host-buffer solve cache (ba.cu): drop the graph gb_ba_solve kept from its previous call on this ctx
void ba_cache_drop(gb_ctx* ctx);
