// Byte movers of the image branch that sit between the cuDNN convolutions.
//
// cp_upsample2x_cat_nhwc: nn.UpsamplingBilinear2d(scale_factor=2) (align_corners=True) applied to the
// channel concatenation torch.cat([img_feat, skip], dim=1) of checkerpose/model/pipeline.py:372-373 and
// :201, in one pass over NHWC tensors: the (B, C1+C2, H, W) concat is never materialised and the
// (B, 2H, 2W, C1+C2) result is written once with 128-bit stores.  HBM-bound: bytes = out + in.
#include "common.cuh"

using bf16 = __nv_bfloat16;

namespace {

struct Src {
  const void* p;
  int64_t sb, sh, sw;  // element strides of batch / row / column; channels are contiguous
  int C;
};

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  __device__ static void load(const bf16* p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
  }
  __device__ static void store(bf16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// Work decomposition: a block of 256 threads produces an 8 x 8 output-pixel patch of one slab of 8 channel vectors
// (64 bf16 / 32 f32 channels) in two passes of 32 pixels.  The patch reads a ~5 x 5 source neighbourhood: every source
// vector is fetched from L2 once and re-read from L1 by the ~4 outputs that use it (a pixel-linear thread order re-reads
// all four taps of every output from L2: 4x the output bytes).
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_cat_nhwc_kernel(Src a, Src b, T* __restrict__ out, int B, int H, int W, float scale_h, float scale_w) {
  constexpr int V = Vec<T>::N;
  const int Ct = a.C + b.C;
  const int cv_per_px = Ct / V;
  const int OH = 2 * H, OW = 2 * W;
  const int slabs = (cv_per_px + 7) >> 3;
  const int cvl = threadIdx.x & 7, pl = threadIdx.x >> 3;   // channel vector in the slab, pixel 0..31 of a pass
  // grid = (slabs * x-tiles, y-tiles, B): the block index IS the work item.  (A grid-stride loop over a linear 64-bit work
  // index spent ~200 of the kernel's 229 instructions per output vector on 64-bit div / mod: ncu r02, 80 % issue slots.)
  {
    const int slab = (int)blockIdx.x % slabs;
    const int bx = (int)blockIdx.x / slabs;
    const int by = (int)blockIdx.y;
    const int bi = (int)blockIdx.z;
    const int cv = slab * 8 + cvl;
    if (cv >= cv_per_px) return;
    int c = cv * V;
    const Src& s = (c < a.C) ? a : b;
    if (c >= a.C) c -= a.C;
    const T* sbase = reinterpret_cast<const T*>(s.p) + (int64_t)bi * s.sb + c;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int p = pass * 32 + pl;
      const int oy = by * 8 + (p >> 3), ox = bx * 8 + (p & 7);
      if (oy >= OH || ox >= OW) continue;
      // same arithmetic as ATen's upsample_bilinear2d with align_corners=True (accumulation type float)
      const float hr = scale_h * (float)oy, wr = scale_w * (float)ox;
      const int h1 = (int)hr, w1 = (int)wr;
      const int h1p = (h1 < H - 1) ? 1 : 0, w1p = (w1 < W - 1) ? 1 : 0;
      const float h1l = hr - (float)h1, h0l = 1.f - h1l;
      const float w1l = wr - (float)w1, w0l = 1.f - w1l;
      const T* base = sbase + (int64_t)h1 * s.sh + (int64_t)w1 * s.sw;
      float v00[V], v01[V], v10[V], v11[V], o[V];
      Vec<T>::load(base, v00);
      Vec<T>::load(base + w1p * s.sw, v01);
      Vec<T>::load(base + h1p * s.sh, v10);
      Vec<T>::load(base + h1p * s.sh + w1p * s.sw, v11);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = h0l * (w0l * v00[i] + w1l * v01[i]) + h1l * (w0l * v10[i] + w1l * v11[i]);
      Vec<T>::store(out + (((int64_t)bi * OH + oy) * OW + ox) * Ct + (int64_t)cv * V, o);
    }
  }
}

}  // namespace

extern "C" int cp_upsample2x_cat_nhwc(const void* a, int64_t a_sb, int64_t a_sh, int64_t a_sw, int Ca, const void* b,
                                      int64_t b_sb, int64_t b_sh, int64_t b_sw, int Cb, int dtype, void* out, int B,
                                      int H, int W, cp_stream_t s) {
  CP_REQUIRE(a && out && B > 0 && H > 0 && W > 0 && Ca > 0 && Cb >= 0, CP_E_INVALID, "cp_upsample2x_cat_nhwc: bad arguments");
  CP_REQUIRE(Cb == 0 || b, CP_E_INVALID, "cp_upsample2x_cat_nhwc: second source is NULL but Cb=%d", Cb);
  const int V = dtype == CP_F32 ? 4 : 8;
  CP_REQUIRE(dtype == CP_F32 || dtype == CP_BF16, CP_E_INVALID, "cp_upsample2x_cat_nhwc: bad dtype %d", dtype);
  CP_REQUIRE(Ca % V == 0 && Cb % V == 0 && a_sb % V == 0 && a_sh % V == 0 && a_sw % V == 0 && b_sb % V == 0 &&
                 b_sh % V == 0 && b_sw % V == 0,
             CP_E_UNSUPPORTED, "cp_upsample2x_cat_nhwc: channels and strides must be multiples of %d elements", V);
  CP_REQUIRE(((uintptr_t)a & 15) == 0 && ((uintptr_t)out & 15) == 0 && (Cb == 0 || ((uintptr_t)b & 15) == 0), CP_E_INVALID,
             "cp_upsample2x_cat_nhwc: pointers must be 16-byte aligned");
  Src sa{a, a_sb, a_sh, a_sw, Ca}, sb{b, b_sb, b_sh, b_sw, Cb};
  // align_corners=True: scale = (in - 1) / (out - 1)
  const float sh = (2 * H > 1) ? (float)(H - 1) / (float)(2 * H - 1) : 0.f;
  const float sw = (2 * W > 1) ? (float)(W - 1) / (float)(2 * W - 1) : 0.f;
  const int cvp = (Ca + Cb) / V;
  CP_REQUIRE(B <= 65535 && (2 * H + 7) / 8 <= 65535, CP_E_UNSUPPORTED, "cp_upsample2x_cat_nhwc: B=%d or H=%d too large for one grid", B, H);
  const dim3 g((unsigned)(((cvp + 7) / 8) * ((2 * W + 7) / 8)), (unsigned)((2 * H + 7) / 8), (unsigned)B);
  if (dtype == CP_F32)
    upsample2x_cat_nhwc_kernel<float><<<g, 256, 0, (cudaStream_t)s>>>(sa, sb, (float*)out, B, H, W, sh, sw);
  else
    upsample2x_cat_nhwc_kernel<bf16><<<g, 256, 0, (cudaStream_t)s>>>(sa, sb, (bf16*)out, B, H, W, sh, sw);
  CP_CHECK_LAUNCH("cp_upsample2x_cat_nhwc");
  return CP_OK;
}
