// gslam_b200/csrc/ba_sweep.cu — the residual + Jacobian sweep of LARGE bundle-adjustment graphs (global BA, BASELINE config 5:
// 500 cameras / 100k landmarks / 1M observations) behind GSLAM::Optimizer::optimize (GSLAM/core/Optimizer.h:229).
//
// Same arithmetic per observation as ba.cu's ba_linearize_kernel (ba_device.cuh: eval_obs / jac_cam / jac_pt; oracle:
// oracle/ba_ref.c), different machine mapping — the local-BA kernel is latency-tuned for 10k observations, this one is
// bandwidth-tuned: the sweep reads 24 B and writes one 144-byte W block per observation (168 B/observation).
//
//   * ONE persistent launch, 2 CTAs of 256 threads per SM; work items are handed out by a device-side ticket (self-resetting:
//     the last CTA to leave zeroes it), heavy items first: [camera slices | landmark groups].
//   * the pose table T_cw (R row-major + t, 96 B per camera) lives in SHARED memory (<= 512 cameras): an observation's pose is a
//     handful of LDS.128 instead of six L1 tag look-ups per lane (every lane of a warp reads a different camera; the old kernel
//     spent 60 % of the L1 data pipe on that).
//   * landmark item = a run of consecutive landmarks with <= 256 observations (host plan, BaDev::lm_goff): ONE LANE PER
//     OBSERVATION (the 8-lanes-per-landmark mapping idled 38 % of the lane rounds at 10 observations per landmark).  The 6x3 W
//     block goes to a shared-memory tile that is contiguous in the global W array, and leaves the SM as ONE bulk asynchronous copy
//     (cp.async.bulk.global.shared::cta, the TMA engine: no LDS/STG instructions, no L1 wavefronts; `UBLKCP` in SASS).  The ten
//     per-observation terms of V_j / g_p,j / cost_j are summed per landmark from shared memory in ascending observation order
//     -> bit-reproducible, and the same order as the oracle.
//   * camera item = one camera (or one slice of a camera with > 8192 observations): 256 threads stride its camera-sorted
//     observation list with the next iteration's indices / point prefetched, 27 accumulators per thread, folded through shared
//     memory in a fixed order (no atomics).  Sliced cameras: last-slice-folds, as in ba.cu.
#include "ba_device.cuh"
#include "ba_internal.cuh"
#include "common.cuh"

using namespace ba;

namespace {

constexpr int kSwThreads = 256;  // threads per CTA = observations per landmark chunk
constexpr int kSwMaxPts = 64;    // landmarks per group (host plan)
constexpr int kSwC = 11;         // doubles per observation in the contribution tile (10 used; odd stride -> conflict-free)
constexpr int kSwPoseCams = 512; // pose table in shared memory up to this many cameras (2 CTAs/SM still fit)

// dynamic shared memory, in doubles
constexpr int kOffW = 0;                               // [256][18]   W tile (bulk-copy source, 16-byte aligned)
constexpr int kOffC = kOffW + kSwThreads * 18;         // [256][11]   per-observation V / g_p / cost terms; camera fold scratch
constexpr int kOffAcc = kOffC + kSwThreads * kSwC;     // [64][10]    per-landmark sums of the group
constexpr int kOffMisc = kOffAcc + kSwMaxPts * 10;     // [40]        pt_off of the group (65 ints), item slots (2 ints), flag
constexpr int kOffPose = kOffMisc + 40;                // [nc][12]    pose table (POSE_SMEM)
static_assert(kSwThreads * 18 + kSwThreads * kSwC >= 27 * kSwThreads, "camera fold scratch spans the W and contribution tiles");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <bool POSE_SMEM>
__device__ __forceinline__ void load_rt(const double* __restrict__ table, int i, double* Rt) {
  const double2* src = reinterpret_cast<const double2*>(table + 12 * (size_t)i);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double2 v = POSE_SMEM ? src[k] : __ldg(src + k);
    Rt[2 * k] = v.x; Rt[2 * k + 1] = v.y;
  }
}

// ---- landmark item: landmarks [j0, j1), observations [pt_off[j0], pt_off[j1]) in chunks of 256 ---------------------------------
template <bool POSE_SMEM>
__device__ __forceinline__ void sweep_landmarks(const BaDev& g, double* sm, const double* __restrict__ pose_tab,
                                                const double* __restrict__ PTS, bool pend, double delta, int grp) {
  const int tid = threadIdx.x;
  double* s_w = sm + kOffW;
  double* s_c = sm + kOffC;
  double* s_acc = sm + kOffAcc;
  int* s_off = reinterpret_cast<int*>(sm + kOffMisc);
  const int j0 = g.lm_goff[grp], j1 = g.lm_goff[grp + 1], L = j1 - j0;
  if (tid <= L) s_off[tid] = g.pt_off[j0 + tid];  // (readers of the previous item are behind the caller's barrier)
  const int e0 = g.pt_off[j0], e1 = g.pt_off[j1];
  for (int c0 = e0; c0 < e1 || c0 == e0; c0 += kSwThreads) {
    if (c0 != e0) {  // the W tile is still being read by the previous chunk's bulk copy; the contribution tile by its item threads
      if (tid == 0) bulk_wait_read();
      __syncthreads();
    }
    const int e = c0 + tid;
    const bool act = e < e1;
    double wv[18], cv[10];
#pragma unroll
    for (int k = 0; k < 18; ++k) wv[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 10; ++k) cv[k] = 0.0;
    if (act) {
      const int i = g.o_cam[e], j = g.o_pt[e];
      const double2 uv = *reinterpret_cast<const double2*>(g.o_uv + 2 * (size_t)e);
      const bool pf = g.pfree[j] != 0;
      const double p[3] = {PTS[3 * (size_t)j], PTS[3 * (size_t)j + 1], PTS[3 * (size_t)j + 2]};
      double Rt[12];
      load_rt<POSE_SMEM>(pose_tab, i, Rt);
      const ObsLin o = eval_obs(Rt, p, uv.x, uv.y, g.has_info ? g.o_info + 3 * (size_t)e : nullptr, delta);
      if (o.valid) {
        cv[9] = o.rho;
        double Jc[12], Jp[6], AJp[6];
        jac_cam(o, g.dof[i], Jc);
        jac_pt(o, Rt, pf, Jp);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          AJp[d] = o.A0 * Jp[d] + o.A1 * Jp[3 + d];
          AJp[3 + d] = o.A1 * Jp[d] + o.A2 * Jp[3 + d];
        }
        const double Ar0 = o.A0 * o.r0 + o.A1 * o.r1, Ar1 = o.A1 * o.r0 + o.A2 * o.r1;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int c = 0; c < 3; ++c) wv[a * 3 + c] = Jc[a] * AJp[c] + Jc[6 + a] * AJp[3 + c];
        int t = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
          for (int c = a; c < 3; ++c) cv[t++] = Jp[a] * AJp[c] + Jp[3 + a] * AJp[3 + c];
          cv[6 + a] = -(Jp[a] * Ar0 + Jp[3 + a] * Ar1);
        }
      }
    }
    {
      double2* mine = reinterpret_cast<double2*>(s_w + 18 * tid);
#pragma unroll
      for (int k = 0; k < 9; ++k) mine[k] = make_double2(wv[2 * k], wv[2 * k + 1]);
#pragma unroll
      for (int k = 0; k < 10; ++k) s_c[kSwC * tid + k] = cv[k];
    }
    fence_async_smem();  // the tile is read by the async proxy next
    __syncthreads();
    const int nobs = min(kSwThreads, e1 - c0);
    if (tid == 0 && nobs > 0) bulk_store(g.W + 18 * (size_t)c0, s_w, (uint32_t)nobs * 144u);
    // per-landmark sums in ascending observation order: thread = (landmark l, term k)
    for (int it = tid; it < L * 10; it += kSwThreads) {
      const int l = it / 10, k = it - l * 10;
      const int a = max(s_off[l], c0) - c0, b = min(s_off[l + 1], c0 + kSwThreads) - c0;
      double s = (c0 == e0) ? 0.0 : s_acc[it];
      for (int o = a; o < b; ++o) s += s_c[kSwC * o + k];
      s_acc[it] = s;
    }
  }
  __syncthreads();
  // V_j, g_p,j, cost_j, the damped inverse, and the installation of an accepted candidate point
  for (int l = tid; l < L; l += kSwThreads) {
    const int j = j0 + l;
    const double* a = s_acc + 10 * l;
    double V[9] = {a[0], a[1], a[2], a[1], a[3], a[4], a[2], a[4], a[5]};
    double* Vg = g.V + 9 * (size_t)j;
#pragma unroll
    for (int k = 0; k < 9; ++k) Vg[k] = V[k];
    g.gp[3 * (size_t)j] = a[6]; g.gp[3 * (size_t)j + 1] = a[7]; g.gp[3 * (size_t)j + 2] = a[8];
    g.cost_pt[j] = a[9];
    if (pend) {
      g.pts[3 * (size_t)j] = PTS[3 * (size_t)j]; g.pts[3 * (size_t)j + 1] = PTS[3 * (size_t)j + 1]; g.pts[3 * (size_t)j + 2] = PTS[3 * (size_t)j + 2];
    }
    const double lambda = g.sc->lambda;
    const bool active = g.pfree[j] != 0 && s_off[l + 1] > s_off[l];
    if (active) {
#pragma unroll
      for (int d = 0; d < 3; ++d) V[d * 4] += lambda * clampd(V[d * 4]);
      if (!spd_inverse<3>(V)) {
#pragma unroll
        for (int k = 0; k < 9; ++k) V[k] = 0.0;
      }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) g.Vinv[9 * (size_t)j + k] = active ? V[k] : 0.0;
  }
}

// ---- camera item: slice `slice` of camera i -----------------------------------------------------------------------------------
template <bool POSE_SMEM>
__device__ __forceinline__ void sweep_camera(const BaDev& g, double* sm, const double* __restrict__ pose_tab,
                                             const double* __restrict__ PTS, bool pend, double delta, int item) {
  const int tid = threadIdx.x;
  const int K = g.cam_split, i = item / K, slice = item - i * K;
  const int dm = g.dof[i];
  const int c0 = g.cam_off[i], c1 = g.cam_off[i + 1];
  const int per = (c1 - c0 + K - 1) / K;
  const int s0 = min(c0 + slice * per, c1), s1 = min(s0 + per, c1);
  double Rt[12];
  load_rt<POSE_SMEM>(pose_tab, i, Rt);
  if (pend && slice == 0) {  // install this camera's accepted pose
    if (tid < 12) g.Rt[12 * i + tid] = g.Rt_new[12 * i + tid];
    else if (tid >= 32 && tid < 39) g.pose[7 * i + tid - 32] = g.pose_new[7 * i + tid - 32];
  }
  double acc[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k] = 0.0;
  int idx = s0 + tid;
  int j = 0;
  double2 uv = make_double2(0.0, 0.0);
  double p[3] = {0.0, 0.0, 0.0};
  if (idx < s1) {
    j = g.c_pt[idx];
    uv = *reinterpret_cast<const double2*>(g.c_uv + 2 * (size_t)idx);
    p[0] = PTS[3 * (size_t)j]; p[1] = PTS[3 * (size_t)j + 1]; p[2] = PTS[3 * (size_t)j + 2];
  }
  while (idx < s1) {
    const int nx = idx + kSwThreads;
    double2 uv_n = make_double2(0.0, 0.0);
    double pn[3] = {0.0, 0.0, 0.0};
    if (nx < s1) {
      const int jn = g.c_pt[nx];
      uv_n = *reinterpret_cast<const double2*>(g.c_uv + 2 * (size_t)nx);
      pn[0] = PTS[3 * (size_t)jn]; pn[1] = PTS[3 * (size_t)jn + 1]; pn[2] = PTS[3 * (size_t)jn + 2];
    }
    const ObsLin o = eval_obs(Rt, p, uv.x, uv.y, g.has_info ? g.o_info + 3 * (size_t)g.cam_perm[idx] : nullptr, delta);
    if (o.valid) {
      double Jc[12], AJc[12];
      jac_cam(o, dm, Jc);
#pragma unroll
      for (int d = 0; d < 6; ++d) {
        AJc[d] = o.A0 * Jc[d] + o.A1 * Jc[6 + d];
        AJc[6 + d] = o.A1 * Jc[d] + o.A2 * Jc[6 + d];
      }
      const double Ar0 = o.A0 * o.r0 + o.A1 * o.r1, Ar1 = o.A1 * o.r0 + o.A2 * o.r1;
      int t = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int b = a; b < 6; ++b) acc[t++] += Jc[a] * AJc[b] + Jc[6 + a] * AJc[6 + b];
      }
#pragma unroll
      for (int a = 0; a < 6; ++a) acc[21 + a] -= Jc[a] * Ar0 + Jc[6 + a] * Ar1;
    }
    idx = nx; uv = uv_n; p[0] = pn[0]; p[1] = pn[1]; p[2] = pn[2];
  }
  // fixed-order fold: term-major scratch [27][256]; warp w folds terms w, w+8, ...: eight strided entries per lane in order, then
  // a fixed shuffle tree
  double* s_red = sm + kOffW;
  double* s_out = sm + kOffAcc;  // [27]
  int* s_flag = reinterpret_cast<int*>(sm + kOffMisc) + 72;
#pragma unroll
  for (int k = 0; k < 27; ++k) s_red[k * kSwThreads + tid] = acc[k];
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  for (int k = warp; k < 27; k += kSwThreads / 32) {
    double r = 0.0;
#pragma unroll
    for (int m = 0; m < kSwThreads / 32; ++m) r += s_red[k * kSwThreads + lane + 32 * m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    if (lane == 0) {
      s_out[k] = r;
      if (K > 1) g.cam_part[((size_t)i * K + slice) * 27 + k] = r;
    }
  }
  if (K > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) *s_flag = (atomicAdd(&g.cam_ticket[i], 1u) == (unsigned)(K - 1)) ? 1 : 0;
    __syncthreads();
    if (!*s_flag) return;
    __threadfence();
    if (tid < 27) {
      double r = 0.0;
      for (int k = 0; k < K; ++k) r += __ldcg(&g.cam_part[((size_t)i * K + k) * 27 + tid]);
      s_out[tid] = r;
    }
    if (tid == 0) g.cam_ticket[i] = 0;
  }
  __syncthreads();
  if (tid < 36) {
    const int a = tid / 6, c = tid % 6, lo = a < c ? a : c, hi = a < c ? c : a;
    const int t = lo * 6 - lo * (lo - 1) / 2 + (hi - lo);
    g.U[36 * i + tid] = s_out[t];
  }
  if (tid < 6) g.gc[6 * i + tid] = s_out[21 + tid];
}

// which: 3 = whole sweep, 1 = camera items only, 2 = landmark items only (timing experiments)
template <bool POSE_SMEM>
__global__ void __launch_bounds__(kSwThreads, 2) ba_sweep_kernel(BaDev g, int which) {
  extern __shared__ __align__(128) double sm[];
  if (g.sc->stop || !g.sc->need_linearize) return;
  const int tid = threadIdx.x;
  const double delta = g.sc->delta;
  const bool pend = g.sc->pending != 0;  // an accepted candidate not installed yet: read the candidate arrays, install on the fly
  const double* PTS = pend ? g.pts_new : g.pts;
  const double* RT = pend ? g.Rt_new : g.Rt;
  const double* pose_tab = RT;
  if (POSE_SMEM) {
    double2* dst = reinterpret_cast<double2*>(sm + kOffPose);
    const double2* src = reinterpret_cast<const double2*>(RT);
    for (int k = tid; k < g.nc * 6; k += kSwThreads) dst[k] = src[k];
    pose_tab = sm + kOffPose;
  }
  const int cam_items = (which & 1) ? g.nc * g.cam_split : 0;
  const int n_items = cam_items + ((which & 2) ? g.lm_ngroups : 0);
  int* s_item = reinterpret_cast<int*>(sm + kOffMisc) + 68;  // two slots: the item being worked on, the one fetched ahead
  if (tid == 0) s_item[0] = (int)atomicAdd(&g.sweep_ticket[0], 1u);
  for (int n = 0;; ++n) {
    if (tid == 0) bulk_wait_read();  // (the W tile of the previous landmark item)
    __syncthreads();
    const int item = s_item[n & 1];
    if (item >= n_items) break;
    int ahead = 0;
    if (tid == 0) ahead = (int)atomicAdd(&g.sweep_ticket[0], 1u);  // consumed after the item: its latency hides behind the work
    if (item < cam_items) sweep_camera<POSE_SMEM>(g, sm, pose_tab, PTS, pend, delta, item);
    else sweep_landmarks<POSE_SMEM>(g, sm, pose_tab, PTS, pend, delta, item - cam_items);
    if (tid == 0) s_item[(n + 1) & 1] = ahead;
  }
  if (tid == 0) {
    bulk_wait_all();
    __threadfence();
    if (atomicAdd(&g.sweep_ticket[1], 1u) == gridDim.x - 1) {  // last CTA out: every CTA has drawn its final (failing) ticket
      g.sweep_ticket[0] = 0u;
      g.sweep_ticket[1] = 0u;
    }
  }
}

}  // namespace

// landmark groups: consecutive landmarks with <= 256 observations and <= 64 landmarks together (a landmark with more observations
// than a chunk is a group of its own, swept in several chunks)
void ba_sweep_plan_host(const std::vector<int>& pt_off, int np, std::vector<int>& goff) {
  goff.clear();
  goff.push_back(0);
  int j = 0;
  while (j < np) {
    int k = j + 1;
    while (k < np && k - j < kSwMaxPts && pt_off[k + 1] - pt_off[j] <= kSwThreads) ++k;
    goff.push_back(k);
    j = k;
  }
}

size_t ba_sweep_smem(const gb_ba_graph* g) {
  const bool pose = g->d.nc <= kSwPoseCams;
  return (size_t)(kOffPose + (pose ? 12 * g->d.nc : 0)) * sizeof(double);
}

int ba_sweep_setup(gb_ctx* ctx) {  // once per device: opt in to the large dynamic shared memory (never lowered)
  static std::once_flag once[64];
  cudaError_t e = cudaSuccess;
  std::call_once(once[ctx->device & 63], [&] {
    const int top = (kOffPose + 12 * kSwPoseCams) * (int)sizeof(double);
    e = cudaFuncSetAttribute(ba_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, top);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ba_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOffPose * (int)sizeof(double));
  });
  if (e != cudaSuccess) { gb_set_error(ctx, "ba_sweep_setup: %s", cudaGetErrorString(e)); return GB_ERR_CUDA; }
  return GB_OK;
}

int ba_sweep_launch(gb_ctx* ctx, gb_ba_graph* g, const BaDev& d, cudaStream_t s, int which) {
  GB_CHECK(ba_sweep_setup(ctx));
  const int items = ((which & 1) ? d.nc * d.cam_split : 0) + ((which & 2) ? d.lm_ngroups : 0);
  if (items <= 0) return GB_OK;
  const bool pose = d.nc <= kSwPoseCams;
  const size_t smem = ba_sweep_smem(g);
  const int grid = std::min(items, ctx->sm_count * 2);
  if (pose) ba_sweep_kernel<true><<<grid, kSwThreads, smem, s>>>(d, which);
  else ba_sweep_kernel<false><<<grid, kSwThreads, smem, s>>>(d, which);
  GB_LAUNCH_CHECK(ctx);
  return GB_OK;
}
