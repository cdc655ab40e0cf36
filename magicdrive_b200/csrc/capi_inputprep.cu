// Input preparation on the device: the step BEFORE the denoising path.
//   mdb_prepare_boxes : magicdrive/dataset/utils.py:120-240 (_preprocess_bbox, bbox_mode "all-xyz", view_shared False,
//                       use_3d_filter True) — per camera view keep the boxes with any corner in front of the camera,
//                       compacted in their original order, 8 corners each, padded to a fixed capacity with masks;
//   mdb_camera_param  : dataset/utils.py:294-297 — [K(3x3) | camera2lidar(3x4)] per view, camera2lidar being the inverse of the
//                       rigid lidar2camera transform ([R^T | -R^T t], demo/helper.py:495-501).
// Boxes are LiDARInstance3DBoxes rows (x, y, z, dx, dy, dz, yaw, ...), bottom-centred; corner order and the z-rotation follow
// mmdet3d as vendored in demo/helper.py:152-190, 39-85.  The visibility test uses the corners of the box re-interpreted with a
// gravity-centre origin (box_center_shift(.., (0.5, 0.5, 0.5)) -> z - dz/2: utils.py / runner/utils.py trans_boxes_to_view),
// while the OUTPUT corners are those of the original box — exactly as the reference does.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/magicdrive_b200.h"
#include "common_host.h"

namespace {

// corner signs in mmdet3d's order (x0y0z0, x0y0z1, x0y1z1, x0y1z0, x1y0z0, x1y0z1, x1y1z1, x1y1z0), origin (0.5, 0.5, 0)
__constant__ float kCx[8] = {-0.5f, -0.5f, -0.5f, -0.5f, 0.5f, 0.5f, 0.5f, 0.5f};
__constant__ float kCy[8] = {-0.5f, -0.5f, 0.5f, 0.5f, -0.5f, -0.5f, 0.5f, 0.5f};
__constant__ float kCz[8] = {0.f, 1.f, 1.f, 0.f, 0.f, 1.f, 1.f, 0.f};

__device__ __forceinline__ void box_corner(const float* b, float zc, int k, float s, float c, float& x, float& y, float& z) {
  const float px = b[3] * kCx[k], py = b[4] * kCy[k], pz = b[5] * kCz[k];
  x = px * c + py * s + b[0];   // points @ rot_mat_T with rot_mat_T = [[cos, -sin, 0], [sin, cos, 0], [0, 0, 1]]
  y = -px * s + py * c + b[1];
  z = pz + zc;
}

__global__ void prepare_boxes_kernel(const float* __restrict__ boxes, int box_dim, const long long* __restrict__ labels,
                                     const int* __restrict__ box_offsets, const float* __restrict__ lidar2camera,
                                     const float* __restrict__ img_aug, int n_views, int capacity, float* __restrict__ out_boxes,
                                     long long* __restrict__ out_classes, uint8_t* __restrict__ out_masks, int* __restrict__ out_counts) {
  const int v = blockIdx.x, s = blockIdx.y;
  const int b0 = box_offsets[s], nb = box_offsets[s + 1] - b0;
  const long long sv = static_cast<long long>(s) * n_views + v;
  __shared__ double zrow[4];  // row 2 of (aug @ lidar2camera): camera-frame depth of a homogeneous lidar point
  __shared__ int warp_cnt[32];
  __shared__ int running;
  if (threadIdx.x < 4) {
    const float* m = lidar2camera + sv * 16;
    double acc;
    if (img_aug) {
      const float* a = img_aug + sv * 16;
      // trans = aug @ lidar2camera in fp32 like the reference (numpy float32 matmul), then used in float64
      float f = 0.f;
      for (int k = 0; k < 4; ++k) f += a[2 * 4 + k] * m[k * 4 + threadIdx.x];
      acc = static_cast<double>(f);
    } else {
      acc = static_cast<double>(m[2 * 4 + threadIdx.x]);
    }
    zrow[threadIdx.x] = acc;
  }
  if (threadIdx.x == 0) running = 0;
  float* ob = out_boxes + sv * capacity * 24;
  long long* oc = out_classes + sv * capacity;
  uint8_t* om = out_masks + sv * capacity;
  for (int i = threadIdx.x; i < capacity; i += blockDim.x) oc[i] = -1, om[i] = 0;
  for (int i = threadIdx.x; i < capacity * 24; i += blockDim.x) ob[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int base = 0; base < nb; base += blockDim.x) {
    const int i = base + threadIdx.x;
    bool vis = false;
    float sn = 0.f, cs = 1.f;
    const float* b = nullptr;
    if (i < nb) {
      b = boxes + static_cast<long long>(b0 + i) * box_dim;
      sn = sinf(b[6]), cs = cosf(b[6]);
      const float zc_shift = b[2] + b[5] * (0.f - 0.5f);  // tensor[:, :3] += dims * (dst - src), dst.z = 0, src.z = 0.5
      for (int k = 0; k < 8; ++k) {
        float x, y, z;
        box_corner(b, zc_shift, k, sn, cs, x, y, z);
        const double zc = zrow[0] * x + zrow[1] * y + zrow[2] * z + zrow[3];
        vis = vis || (zc > 0.0);
      }
    }
    const unsigned ball = __ballot_sync(0xffffffffu, vis);
    if (lane == 0) warp_cnt[warp] = __popc(ball);
    __syncthreads();
    int before = running;
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    const int pos = before + __popc(ball & ((1u << lane) - 1u));
    if (vis && pos < capacity) {
      for (int k = 0; k < 8; ++k) {
        float x, y, z;
        box_corner(b, b[2], k, sn, cs, x, y, z);
        ob[pos * 24 + k * 3 + 0] = x, ob[pos * 24 + k * 3 + 1] = y, ob[pos * 24 + k * 3 + 2] = z;
      }
      oc[pos] = labels[b0 + i];
      om[pos] = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = running;
      for (int w = 0; w < nwarps; ++w) t += warp_cnt[w];
      running = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_counts[sv] = running;  // may exceed `capacity`: the caller checks
}

__global__ void camera_param_kernel(const float* __restrict__ intrinsics, const float* __restrict__ lidar2camera, int n,
                                    float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* K = intrinsics + static_cast<long long>(i) * 16;
  const float* M = lidar2camera + static_cast<long long>(i) * 16;
  float* o = out + static_cast<long long>(i) * 21;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) o[r * 7 + c] = K[r * 4 + c];
    for (int c = 0; c < 3; ++c) o[r * 7 + 3 + c] = M[c * 4 + r];  // R^T
    float t = 0.f;
    for (int k = 0; k < 3; ++k) t += -M[k * 4 + r] * M[k * 4 + 3];  // bmm(-R^T, t)
    o[r * 7 + 6] = t;
  }
}

}  // namespace

extern "C" int mdb_prepare_boxes(const float* boxes, int box_dim, const long long* labels, const int* box_offsets, int n_scenes,
                                 const float* lidar2camera, const float* img_aug, int n_views, int capacity, float* out_boxes,
                                 long long* out_classes, unsigned char* out_masks, int* out_counts, void* stream) {
  using namespace mdb;
  if (!boxes || !labels || !box_offsets || !lidar2camera || !out_boxes || !out_classes || !out_masks || !out_counts)
    return set_error(MDB_ERR_INVALID, "mdb_prepare_boxes: null pointer");
  if (box_dim < 7 || n_scenes <= 0 || n_views <= 0 || capacity <= 0)
    return set_error(MDB_ERR_INVALID, "mdb_prepare_boxes: bad sizes (box_dim=%d scenes=%d views=%d capacity=%d)", box_dim, n_scenes,
                     n_views, capacity);
  prepare_boxes_kernel<<<dim3(n_views, n_scenes), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      boxes, box_dim, labels, box_offsets, lidar2camera, img_aug, n_views, capacity, out_boxes, out_classes, out_masks, out_counts);
  MDB_CHECK_LAUNCH("prepare_boxes_kernel");
  return MDB_OK;
}

extern "C" int mdb_camera_param(const float* intrinsics, const float* lidar2camera, int n, float* out, void* stream) {
  using namespace mdb;
  if (!intrinsics || !lidar2camera || !out || n <= 0) return set_error(MDB_ERR_INVALID, "mdb_camera_param: bad arguments");
  camera_param_kernel<<<(n + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(intrinsics, lidar2camera, n, out);
  MDB_CHECK_LAUNCH("camera_param_kernel");
  return MDB_OK;
}
