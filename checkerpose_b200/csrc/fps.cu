// Farthest point sampling of a CAD model's vertices: the offline step that produces the keypoint files the head is
// built from (checkerpose/preprocess_data/get_fps_points.py:65-90, farthest_point_sample_init_center).  SURVEY.md
// section 8(f) rank 4: not on the per-RoI path, here for the "keypoints in -> codes out" surface.
//
// float64 throughout and the reference's operation order (numpy: d = sqrt((dx*dx + dy*dy) + dz*dz), no fused
// multiply-add; `distances < distances_to_set` update; np.argmax = first maximum), so that the ids are bit-exact.
// One CTA of 1024 threads walks all vertices once per sample; the argmax is a warp-shuffle + shared-memory reduction
// on (distance, lowest index).  O(npoint * V) like the reference (whose NumPy loop takes minutes per object).
#include "common.cuh"

namespace {

constexpr int FPS_THREADS = 1024;

__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_kernel(const double* __restrict__ xyz, int V, int npoint, double cx, double cy, double cz, double init_dist,
           double* __restrict__ dist, int64_t* __restrict__ ids, double* __restrict__ out_xyz) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_far;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < V; i += FPS_THREADS) dist[i] = init_dist;
  double fx = cx, fy = cy, fz = cz;
  __syncthreads();
  for (int sidx = 0; sidx < npoint; ++sidx) {
    double best = -1.0;
    int besti = 0x7fffffff;
    for (int i = tid; i < V; i += FPS_THREADS) {
      const double dx = xyz[3 * i] - fx, dy = xyz[3 * i + 1] - fy, dz = xyz[3 * i + 2] - fz;
      const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
      double ds = dist[i];
      if (d < ds) {
        ds = d;
        dist[i] = d;
      }
      if (ds > best) {      // strictly greater: the first (lowest-index) maximum of this thread's strided points
        best = ds;
        besti = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) {
        best = ov;
        besti = oi;
      }
    }
    if (lane == 0) {
      s_val[warp] = best;
      s_idx[warp] = besti;
    }
    __syncthreads();
    if (warp == 0) {
      best = s_val[lane];
      besti = s_idx[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) {
          best = ov;
          besti = oi;
        }
      }
      if (lane == 0) s_far = besti;
    }
    __syncthreads();
    const int far = s_far;
    fx = xyz[3 * far];
    fy = xyz[3 * far + 1];
    fz = xyz[3 * far + 2];
    if (tid == 0) {
      ids[sidx] = far;
      out_xyz[3 * sidx] = fx;
      out_xyz[3 * sidx + 1] = fy;
      out_xyz[3 * sidx + 2] = fz;
    }
    __syncthreads();   // s_val / s_idx / s_far are rewritten by the next sample
  }
}

}  // namespace

extern "C" int cp_fps(const double* xyz, int V, int npoint, const double* center, double init_dist, double* dist_ws,
                      int64_t* ids, double* fps_xyz, cp_stream_t s) {
  CP_REQUIRE(xyz && center && dist_ws && ids && fps_xyz, CP_E_INVALID, "cp_fps: null pointer");
  CP_REQUIRE(V > 0 && npoint > 0, CP_E_INVALID, "cp_fps: bad sizes V=%d npoint=%d", V, npoint);
  fps_kernel<<<1, FPS_THREADS, 0, (cudaStream_t)s>>>(xyz, V, npoint, center[0], center[1], center[2], init_dist, dist_ws, ids, fps_xyz);
  CP_CHECK_LAUNCH("cp_fps");
  return CP_OK;
}
