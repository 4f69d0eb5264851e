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

// Work decomposition: a block of 256 threads produces a 16 x 16 output-pixel patch of one slab of 8 channel vectors
// (64 bf16 / 32 f32 channels) in eight passes of 32 pixels (4 rows x 8 columns each); grid = (slabs * x-tiles, y-tiles, B).
// The patch reads a ~9 x 9 source neighbourhood: every source vector is fetched from L2 once and re-read from L1 by the ~4
// outputs that use it; the per-thread index arithmetic is amortised over eight output vectors.
// The kernel is issue-bound, not HBM-bound (ncu r02: 80 % issue slots): element offsets inside a RoI are 32-bit, the four
// bilinear weights are combined once per pixel (4 FMAs per element instead of 6 ops; within 1 ulp of ATen's grouping
// h0*(w0*a + w1*b) + h1*(w0*c + w1*d), then rounded to the output type), bf16 pairs are widened with one shift / one mask.
template <typename T> struct Wide;
template <> struct Wide<float> {
  using Raw = float4;
  __device__ static Raw load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ static void fma(float (&o)[4], float w, const Raw& v, bool first) {
    if (first) { o[0] = w * v.x; o[1] = w * v.y; o[2] = w * v.z; o[3] = w * v.w; }
    else { o[0] = fmaf(w, v.x, o[0]); o[1] = fmaf(w, v.y, o[1]); o[2] = fmaf(w, v.z, o[2]); o[3] = fmaf(w, v.w, o[3]); }
  }
};
template <> struct Wide<bf16> {
  using Raw = uint4;
  __device__ static Raw load(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static void fma(float (&o)[8], float w, const Raw& v, bool first) {
    const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lo = __uint_as_float(ww[i] << 16), hi = __uint_as_float(ww[i] & 0xffff0000u);
      o[2 * i] = first ? w * lo : fmaf(w, lo, o[2 * i]);
      o[2 * i + 1] = first ? w * hi : fmaf(w, hi, o[2 * i + 1]);
    }
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_cat_nhwc_kernel(Src a, Src b, T* __restrict__ out, int64_t o_sb, int o_sh, int o_sw, int B, int H, int W, float scale_h, float scale_w) {
  constexpr int V = Vec<T>::N;
  const int Ct = a.C + b.C;
  const int cv_per_px = Ct / V;
  const int OH = 2 * H, OW = 2 * W;
  const int slabs = (cv_per_px + 7) >> 3;
  const int cvl = threadIdx.x & 7, pl = threadIdx.x >> 3;   // channel vector in the slab, pixel 0..31 of a pass
  const int slab = (int)blockIdx.x % slabs;
  const int bx = (int)blockIdx.x / slabs;
  const int by = (int)blockIdx.y;
  const int bi = (int)blockIdx.z;
  const int cv = slab * 8 + cvl;
  if (cv >= cv_per_px) return;
  int c = cv * V;
  const bool from_a = c < a.C;
  if (!from_a) c -= a.C;
  const int64_t s_sb = from_a ? a.sb : b.sb;
  const int s_sh = (int)(from_a ? a.sh : b.sh), s_sw = (int)(from_a ? a.sw : b.sw);      // < 2^31: checked by the host
  const T* sbase = reinterpret_cast<const T*>(from_a ? a.p : b.p) + (int64_t)bi * s_sb + c;
  T* obase = out + (int64_t)bi * o_sb + cv * V;
#pragma unroll 2
  for (int pass = 0; pass < 8; ++pass) {
    // pass -> 4 x 8 pixel block (pass >> 1 = block row, pass & 1 = block column) of the 16 x 16 patch
    const int oy = by * 16 + (pass >> 1) * 4 + (pl >> 3), ox = bx * 16 + (pass & 1) * 8 + (pl & 7);
    if (oy >= OH || ox >= OW) continue;
    // source coordinates as ATen's upsample_bilinear2d with align_corners=True (accumulation type float)
    const float hr = scale_h * (float)oy, wr = scale_w * (float)ox;
    const int h1 = (int)hr, w1 = (int)wr;
    const int dh = (h1 < H - 1) ? s_sh : 0, dw = (w1 < W - 1) ? s_sw : 0;
    const float h1l = hr - (float)h1, h0l = 1.f - h1l;
    const float w1l = wr - (float)w1, w0l = 1.f - w1l;
    const int o00 = h1 * s_sh + w1 * s_sw;
    const typename Wide<T>::Raw v00 = Wide<T>::load(sbase + o00), v01 = Wide<T>::load(sbase + o00 + dw);
    const typename Wide<T>::Raw v10 = Wide<T>::load(sbase + o00 + dh), v11 = Wide<T>::load(sbase + o00 + dh + dw);
    float o[V];
    Wide<T>::fma(o, h0l * w0l, v00, true);
    Wide<T>::fma(o, h0l * w1l, v01, false);
    Wide<T>::fma(o, h1l * w0l, v10, false);
    Wide<T>::fma(o, h1l * w1l, v11, false);
    Vec<T>::store(obase + oy * o_sh + ox * o_sw, o);
  }
}

}  // namespace

extern "C" int cp_upsample2x_cat_nhwc_to(const void* a, int64_t a_sb, int64_t a_sh, int64_t a_sw, int Ca, const void* b,
                                         int64_t b_sb, int64_t b_sh, int64_t b_sw, int Cb, int dtype, void* out, int64_t o_sb,
                                         int64_t o_sh, int64_t o_sw, int B, int H, int W, cp_stream_t s) {
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
  CP_REQUIRE((int64_t)(H - 1) * a_sh + (int64_t)(W - 1) * a_sw < (1ll << 31) && (int64_t)(H - 1) * b_sh + (int64_t)(W - 1) * b_sw < (1ll << 31) &&
             (int64_t)(2 * H - 1) * o_sh + (int64_t)(2 * W - 1) * o_sw + Ca + Cb < (1ll << 31) && o_sb % V == 0 && o_sh % V == 0 && o_sw % V == 0 &&
             o_sw >= Ca + Cb, CP_E_UNSUPPORTED, "cp_upsample2x_cat_nhwc: one RoI's map must stay below 2^31 elements");
  CP_REQUIRE(B <= 65535 && (2 * H + 15) / 16 <= 65535, CP_E_UNSUPPORTED, "cp_upsample2x_cat_nhwc: B=%d or H=%d too large for one grid", B, H);
  const dim3 g((unsigned)(((cvp + 7) / 8) * ((2 * W + 15) / 16)), (unsigned)((2 * H + 15) / 16), (unsigned)B);
  if (dtype == CP_F32)
    upsample2x_cat_nhwc_kernel<float><<<g, 256, 0, (cudaStream_t)s>>>(sa, sb, (float*)out, o_sb, (int)o_sh, (int)o_sw, B, H, W, sh, sw);
  else
    upsample2x_cat_nhwc_kernel<bf16><<<g, 256, 0, (cudaStream_t)s>>>(sa, sb, (bf16*)out, o_sb, (int)o_sh, (int)o_sw, B, H, W, sh, sw);
  CP_CHECK_LAUNCH("cp_upsample2x_cat_nhwc");
  return CP_OK;
}

extern "C" int cp_upsample2x_cat_nhwc(const void* a, int64_t a_sb, int64_t a_sh, int64_t a_sw, int Ca, const void* b,
                                      int64_t b_sb, int64_t b_sh, int64_t b_sw, int Cb, int dtype, void* out, int B,
                                      int H, int W, cp_stream_t s) {
  const int64_t Ct = (int64_t)Ca + Cb;
  return cp_upsample2x_cat_nhwc_to(a, a_sb, a_sh, a_sw, Ca, b, b_sb, b_sh, b_sw, Cb, dtype, out, 4 * (int64_t)H * W * Ct, 2 * (int64_t)W * Ct, Ct,
                                   B, H, W, s);
}

// ---- zero border of a bordered NHWC map (the slab convolutions read their padding from the map itself): the LAST row and
//      the LAST column of every image ----
namespace {
__global__ void zero_border_kernel(uint4* __restrict__ x, int Hp, int Wp, int row16) {
  const int nb = Wp + Hp - 1;       // border pixels per image
  const int e = (int)blockIdx.x % nb, b = (int)blockIdx.x / nb;
  const int py = e < Wp ? Hp - 1 : e - Wp, px = e < Wp ? e : Wp - 1;
  uint4* row = x + (((int64_t)b * Hp + py) * Wp + px) * row16;
  for (int i = threadIdx.x; i < row16; i += blockDim.x) row[i] = make_uint4(0, 0, 0, 0);
}
}  // namespace

extern "C" int cp_zero_border_nhwc(void* x, int B, int Hp, int Wp, int C, int elem_bytes, cp_stream_t s) {
  CP_REQUIRE(x && B > 0 && Hp >= 2 && Wp >= 2 && C > 0 && (elem_bytes == 2 || elem_bytes == 4), CP_E_INVALID, "cp_zero_border_nhwc: bad arguments");
  CP_REQUIRE(((int64_t)C * elem_bytes) % 16 == 0 && ((uintptr_t)x & 15) == 0, CP_E_UNSUPPORTED, "cp_zero_border_nhwc: pixel rows must be whole 16-byte vectors");
  const int64_t blocks = (int64_t)B * (Wp + Hp - 1);
  CP_REQUIRE(blocks < (1ll << 31), CP_E_UNSUPPORTED, "cp_zero_border_nhwc: too many border pixels");
  const int row16 = (int)((int64_t)C * elem_bytes / 16);
  zero_border_kernel<<<(unsigned)blocks, row16 < 128 ? 32 : 64, 0, (cudaStream_t)s>>>((uint4*)x, Hp, Wp, row16);
  CP_CHECK_LAUNCH("cp_zero_border_nhwc");
  return CP_OK;
}
