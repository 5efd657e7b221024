// Sampled motion -> 3-D joints, fused into one launch: de-normalisation, yaw / root-trajectory integration, local-joint
// rotation and the pair's init pose.  Last stage of "caption -> joints"  and the basis of the
// MPJPE parity metric.
//
// Reference: tools/visualization.py:149-155                 motion[1:] * std + mean, motion[0,:4] * init_std + init_mean,
//                                                           init-state row moved to the end
//            utils/motion_process.py:362-381                recover_root_rot_pos (two cumsums, qrot(qinv(q), v))
//            utils/motion_process.py:418-456                recover_from_ric2
//            utils/quaternion.py:16-20, 54-73               qinv, qrot
//
// Every quaternion on this path is (w, 0, y, 0), so qrot reduces to a planar rotation; it is evaluated in the reference's
// association order with unfused multiplies/adds:  uv = u x v,  uuv = u x uv,  out = v + 2 (w uv + uuv).
// torch's CPU cumsum accumulates float32 in float64 and rounds each prefix to float32; the two prefix sums here do the
// same, sequentially (T <= 196 in the product: ~600 dependent DADDs per sequence, off the critical path), so the only
// differences left against the CPU reference are the last-ulp of cosf / sinf.
//
// One CTA per sequence.  HBM traffic: reads the 67 used columns of every row (the sectors of 268 B out of 1052 B), writes
// 264 B per frame: S = 128, T = 196 -> 6.7 MB in + 6.6 MB out; latency-bound by the two scans, ~10 us per launch.
#include "hig_common.cuh"
#include "hig_internal.h"

namespace hig {

constexpr int JT_THREADS = 256;
constexpr int JT_MAX_T = 2048;

// rotation of (vx, vy, vz) by the quaternion (w, 0, u, 0): y is untouched
__device__ __forceinline__ void qrot_y(float w, float u, float vx, float vz, float& ox, float& oz) {
  const float uv_x = __fmul_rn(u, vz);
  const float uv_z = -__fmul_rn(u, vx);
  const float uuv_x = __fmul_rn(u, uv_z);
  const float uuv_z = -__fmul_rn(u, uv_x);
  ox = __fadd_rn(vx, __fmul_rn(2.f, __fadd_rn(__fmul_rn(w, uv_x), uuv_x)));
  oz = __fadd_rn(vz, __fmul_rn(2.f, __fadd_rn(__fmul_rn(w, uv_z), uuv_z)));
}

__global__ void __launch_bounds__(JT_THREADS)
recover_joints_kernel(const float* __restrict__ x, int T, int C, int init_row, const float* __restrict__ mean,
                      const float* __restrict__ stdv, const float* __restrict__ init_mean,
                      const float* __restrict__ init_std, const int* __restrict__ length, int J,
                      float* __restrict__ joints) {
  extern __shared__ float sm[];
  const int F = T - 1;                       // motion frames: every row but the init-state row
  float* s_c = sm;                           // cos(yaw)            (first: yaw velocity)
  float* s_s = sm + F;                       // sin(yaw)
  float* s_x = sm + 2 * F;                   // root x              (first: rotated x velocity)
  float* s_z = sm + 3 * F;                   // root z
  float* s_v = sm + 4 * F;                   // raw XZ velocity, 2 per frame
  const int s = blockIdx.x;
  const float* xs = x + (size_t)s * T * C;
  const int first = init_row == 0 ? 1 : 0;   // row of frame 0
  const bool norm = mean != nullptr;
  auto feat = [&](int f, int c) {            // de-normalised feature c of frame f (:149-150)
    const float v = xs[(size_t)(f + first) * C + c];
    return norm ? __fadd_rn(__fmul_rn(v, stdv[c]), mean[c]) : v;
  };

  for (int f = threadIdx.x; f < F; f += JT_THREADS) {
    s_c[f] = feat(f, 0);
    s_v[2 * f] = feat(f, 1);
    s_v[2 * f + 1] = feat(f, 2);
  }
  __syncthreads();
  if (threadIdx.x == 0) {                    // r_rot_ang = cumsum([0, rot_vel[:-1]])  (:365-368)
    double acc = 0.0;
    float prev = s_c[0];
    float ang = 0.f;
    for (int f = 0; f < F; ++f) {
      if (f > 0) { acc += (double)prev; ang = (float)acc; }
      prev = s_c[f];
      s_c[f] = cosf(ang);
      s_s[f] = sinf(ang);
    }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += JT_THREADS) {   // r_pos[1:, (0,2)] = vel[:-1]; qrot(qinv(q), r_pos)  (:374-377)
    float ox = 0.f, oz = 0.f;
    if (f > 0) qrot_y(s_c[f], -s_s[f], s_v[2 * (f - 1)], s_v[2 * (f - 1) + 1], ox, oz);
    else qrot_y(s_c[0], -s_s[0], 0.f, 0.f, ox, oz);
    s_x[f] = ox;
    s_z[f] = oz;
  }
  __syncthreads();
  if (threadIdx.x < 2) {                     // r_pos = cumsum(r_pos, dim=-2)  (:379)
    float* p = threadIdx.x == 0 ? s_x : s_z;
    double acc = 0.0;
    for (int f = 0; f < F; ++f) { acc += (double)p[f]; p[f] = (float)acc; }
  }
  __syncthreads();

  // init state (:419-425, :151-152): position (x, z) and the un-normalised quaternion (w, 0, y, 0)
  float ini[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float v = xs[(size_t)init_row * C + i];
    ini[i] = init_mean != nullptr ? __fadd_rn(__fmul_rn(v, init_std[i]), init_mean[i]) : v;
  }
  const int valid = length != nullptr ? max(0, min(F, length[s] - 1)) : F;
  float* out = joints + (size_t)s * F * J * 3;
  for (int idx = threadIdx.x; idx < F * J; idx += JT_THREADS) {
    const int f = idx / J, j = idx - f * J;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (f < valid) {
      if (j == 0) {                          // root: (r_pos.x, data[..., 3], r_pos.z)  (:381, :447)
        px = s_x[f]; py = feat(f, 3); pz = s_z[f];
      } else {                               // qrot(qinv(q), local) + root XZ  (:436-444)
        const int c = 4 + 3 * (j - 1);
        float rx, rz;
        qrot_y(s_c[f], -s_s[f], feat(f, c), feat(f, c + 2), rx, rz);
        px = __fadd_rn(rx, s_x[f]);
        py = feat(f, c + 1);
        pz = __fadd_rn(rz, s_z[f]);
      }
      float gx, gz;                          // qrot(root_quat_init, p) + init_pos  (:450-456)
      qrot_y(ini[2], ini[3], px, pz, gx, gz);
      px = __fadd_rn(gx, ini[0]);
      pz = __fadd_rn(gz, ini[1]);
    }
    out[(size_t)idx * 3 + 0] = px;
    out[(size_t)idx * 3 + 1] = py;
    out[(size_t)idx * 3 + 2] = pz;
  }
}

int recover_joints(const float* x, int S, int T, int C, int init_row, const float* mean, const float* stdv,
                   const float* init_mean, const float* init_std, const int* length, int joints_num, float* joints,
                   cudaStream_t stream) {
  if (!x || !joints || S <= 0) return set_error(HIG_ERR_INVALID, "recover_joints: null operand or empty batch");
  if (T < 2 || T > JT_MAX_T) return set_error(HIG_ERR_UNSUPPORTED, "recover_joints: 2 <= T <= 2048 rows (init state + frames)");
  if (joints_num < 1 || C < 4 + 3 * (joints_num - 1)) return set_error(HIG_ERR_INVALID, "recover_joints: feature width too small for joints_num");
  if (init_row != 0 && init_row != T - 1) return set_error(HIG_ERR_INVALID, "recover_joints: the init-state row is the first or the last row");
  if ((mean == nullptr) != (stdv == nullptr) || (init_mean == nullptr) != (init_std == nullptr))
    return set_error(HIG_ERR_INVALID, "recover_joints: mean/std (and init_mean/init_std) come in pairs");
  const size_t smem = (size_t)6 * (T - 1) * sizeof(float);
  recover_joints_kernel<<<S, JT_THREADS, smem, stream>>>(x, T, C, init_row, mean, stdv, init_mean, init_std, length,
                                                        joints_num, joints);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HIG_ERR_CUDA, std::string("recover_joints launch: ") + cudaGetErrorString(e));
  count_launch();
  return HIG_OK;
}

}  // namespace hig
