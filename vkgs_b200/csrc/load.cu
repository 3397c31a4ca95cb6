// Load-time activation: raw PLY floats -> resident scene.  Replaces parse_ply.comp:34-98 (dispatch
// engine.cc:1137-1152): s = exp(scale), q = (rot_1,rot_2,rot_3,rot_0)/|q| -> R, Sigma = R diag(s^2) R^T,
// SH f32 -> f16 (RNE), opacity = sigmoid.  Same 60-entry float-offset table as splat_load_thread.cc:114-135.
// Output layout is the B200 one (planar positions + one 128-byte payload line per splat, common.cuh), not the
// reference's four SoA buffers; k_export_scene converts back for the parity tap.
//
// Compiled with -fmad=false so the covariance arithmetic is the oracle's operation for operation; exp() is the
// only non-reproducible ingredient (device expf vs libm: <= 2 ulp), hence activation parity is tolerance-checked.
#include "common.cuh"
#include "kernels.h"

namespace vkgsb {

__device__ __forceinline__ void mat3_mul_l(const float* A, const float* B, float* C) {
  float t[9];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r)
      t[c * 3 + r] = (A[0 * 3 + r] * B[c * 3 + 0] + A[1 * 3 + r] * B[c * 3 + 1]) + A[2 * 3 + r] * B[c * 3 + 2];
#pragma unroll
  for (int i = 0; i < 9; ++i) C[i] = t[i];
}

__global__ void __launch_bounds__(256)
k_activate(const float* __restrict__ rows, const uint32_t* __restrict__ offsets, uint32_t first, uint32_t count,
           SceneStorage dst) {
  __shared__ uint32_t off[60];
  if (threadIdx.x < 60) off[threadIdx.x] = offsets[threadIdx.x];
  __syncthreads();
  const uint32_t i = blockIdx.x * 256 + threadIdx.x;
  if (i >= count) return;
  const float* p = rows + static_cast<size_t>(off[59]) * i;
  const uint32_t id = first + i;

  const float s0 = expf(p[off[3]]), s1 = expf(p[off[4]]), s2 = expf(p[off[5]]);
  float qx = p[off[6]], qy = p[off[7]], qz = p[off[8]], qw = p[off[9]];
  const float len = sqrtf(((qx * qx + qy * qy) + qz * qz) + qw * qw);
  qx = qx / len; qy = qy / len; qz = qz / len; qw = qw / len;
  const float xx = qx * qx, yy = qy * qy, zz = qz * qz, xy = qx * qy, xz = qx * qz, yz = qy * qz;
  const float wx = qw * qx, wy = qw * qy, wz = qw * qz;
  float rot[9];
  rot[0] = 1.f - 2.f * (yy + zz); rot[1] = 2.f * (xy + wz);       rot[2] = 2.f * (xz - wy);
  rot[3] = 2.f * (xy - wz);       rot[4] = 1.f - 2.f * (xx + zz); rot[5] = 2.f * (yz + wx);
  rot[6] = 2.f * (xz + wy);       rot[7] = 2.f * (yz - wx);       rot[8] = 1.f - 2.f * (xx + yy);
  float ss[9] = {s0 * s0, 0.f, 0.f, 0.f, s1 * s1, 0.f, 0.f, 0.f, s2 * s2};
  float rt[9], c3[9];
  mat3_mul_l(rot, ss, c3);
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 3; ++r) rt[c * 3 + r] = rot[r * 3 + c];
  mat3_mul_l(c3, rt, c3);

  dst.x[id] = p[off[0]];
  dst.y[id] = p[off[1]];
  dst.z[id] = p[off[2]];
  // the largest eigenvalue of Sigma = R diag(s^2) R^T: bounds every projected extent (band cull, project.cu)
  dst.tr[id] = fmaxf(fmaxf(s0 * s0, s1 * s1), s2 * s2);

  uint4 line[8];
  line[0] = make_uint4(__float_as_uint(c3[0]), __float_as_uint(c3[3]), __float_as_uint(c3[6]), __float_as_uint(c3[4]));
  line[1] = make_uint4(__float_as_uint(c3[7]), __float_as_uint(c3[8]),
                       __float_as_uint(1.f / (1.f + expf(-p[off[58]]))), 0u);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half lo = __float2half_rn(p[off[10 + 8 * k + 2 * j]]), hi = __float2half_rn(p[off[10 + 8 * k + 2 * j + 1]]);
      w[j] = static_cast<uint32_t>(__half_as_ushort(lo)) | (static_cast<uint32_t>(__half_as_ushort(hi)) << 16);
    }
    line[2 + k] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  uint4* out = reinterpret_cast<uint4*>(dst.payload + id);
#pragma unroll
  for (int k = 0; k < 8; ++k) out[k] = line[k];
}

__global__ void __launch_bounds__(256)
k_export_scene(SceneStorage src, uint32_t n, float* __restrict__ pos, float* __restrict__ cov,
               float* __restrict__ opacity, uint16_t* __restrict__ sh) {
  const uint32_t id = blockIdx.x * 256 + threadIdx.x;
  if (id >= n) return;
  const SplatPayload& pl = src.payload[id];
  if (pos) {
    pos[3 * static_cast<size_t>(id) + 0] = src.x[id];
    pos[3 * static_cast<size_t>(id) + 1] = src.y[id];
    pos[3 * static_cast<size_t>(id) + 2] = src.z[id];
  }
  if (cov)
    for (int k = 0; k < 6; ++k) cov[6 * static_cast<size_t>(id) + k] = pl.cov[k];
  if (opacity) opacity[id] = pl.opacity;
  if (sh)
    for (int k = 0; k < 48; ++k) sh[48 * static_cast<size_t>(id) + k] = __half_as_ushort(pl.sh[k]);
}

void launch_activate(const float* d_rows, const uint32_t* d_offsets, uint32_t first, uint32_t count,
                     const SceneStorage& dst, cudaStream_t stream) {
  if (count == 0) return;
  k_activate<<<(count + 255) / 256, 256, 0, stream>>>(d_rows, d_offsets, first, count, dst);
}

void launch_export_scene(const SceneStorage& src, uint32_t n, float* d_pos, float* d_cov, float* d_opacity,
                         uint16_t* d_sh, cudaStream_t stream) {
  if (n == 0) return;
  k_export_scene<<<(n + 255) / 256, 256, 0, stream>>>(src, n, d_pos, d_cov, d_opacity, d_sh);
}

}  // namespace vkgsb
