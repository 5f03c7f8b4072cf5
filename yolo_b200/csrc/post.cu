// Post-decode consumers that sit between predict() and the ROS publish on every frame:
//   azimuth      car/video_node.py:244-252, yolo_modules/yolo_cv.py:85-94 (`cls2ang`): softmax over the 24 orientation-class logits,
//                circular mean  atan2(sum sin_k p_k, sum cos_k p_k)  with k * 360/24 degree offsets, radius = confidence * |mean vector|
//   plate corners  yolo_modules/licence_plate_render/__init__.py:340-377 (`ProjectRectangle6D.__call__` / `projection_matrix`):
//                the four corners of the 399 x 168 mm plate under the predicted 6-D pose, projected with the camera intrinsics
//   plate un-warp  `add_edges` :379-402: cv2.getPerspectiveTransform(corners -> 380 x 160 rectangle) + cv2.warpPerspective (bilinear,
//                constant border) of the camera frame - the crop the OCR stage reads
// All three are tiny; they exist so that the per-frame path never leaves the device between the network and its consumers.
#include <math.h>

#include "common.cuh"

namespace yb {

__global__ void azimuth_kernel(const float* __restrict__ rows, int batch, int row_len, int n_class, float* __restrict__ out_ang,
                               float* __restrict__ out_rad) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float* x = rows + (size_t)b * row_len + (row_len - n_class);
  float mx = x[0];
  for (int k = 1; k < n_class; ++k) mx = fmaxf(mx, x[k]);
  double se = 0.0, c = 0.0, s = 0.0;
  for (int k = 0; k < n_class; ++k) {
    const double e = exp((double)x[k] - (double)mx);
    const double a = (double)(k * (360 / n_class)) * 3.14159265358979323846 / 180.0;     // range(0, 360, 360 / 24), car/video_node.py:36-38
    se += e; c += cos(a) * e; s += sin(a) * e;
  }
  c /= se; s /= se;
  out_ang[b] = (float)atan2(s, c);
  if (out_rad) out_rad[b] = (float)((double)rows[(size_t)b * row_len] * sqrt(s * s + c * c));
}

// corners[b][i] = (u, v) of plate corner i in camera pixels (x then y), times (x_scale, y_scale)
__global__ void lp_corners_kernel(const float* __restrict__ poses, int batch, int pose_stride, int pose_off, double fx, double fy, double cx, double cy,
                                  float x_scale, float y_scale, float* __restrict__ corners) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float* q = poses + (size_t)b * pose_stride + pose_off;
  const double X = q[0], Y = q[1], Z = q[2], r1 = q[3], r2 = q[4], r3 = q[5];
  const double a = sin(r1) * cos(r2) * 84.0, bb = sin(r1) * sin(r2) * cos(r3) * 84.0, c = sin(r2) * 199.5, d = sin(r3) * cos(r1) * 84.0;
  const double e = cos(r2) * cos(r3) * 199.5, f = sin(r1) * sin(r2) * sin(r3) * 84.0, g = sin(r3) * cos(r2) * 199.5, h = cos(r1) * cos(r3) * 84.0;
  const double zz[4] = {Z + a - c, Z + a + c, Z - a + c, Z - a - c};
  const double xx[4] = {X + bb - d + e, X + bb - d - e, X - bb + d - e, X - bb + d + e};
  const double yy[4] = {Y + f + g + h, Y + f - g + h, Y - f - g - h, Y - f + g - h};
  for (int i = 0; i < 4; ++i) {
    const float u = (float)((cx * zz[i] + fx * xx[i]) / zz[i]), v = (float)((cy * zz[i] + fy * yy[i]) / zz[i]);   // points.astype(np.float32)
    corners[((size_t)b * 4 + i) * 2] = u * x_scale;
    corners[((size_t)b * 4 + i) * 2 + 1] = v * y_scale;
  }
}

// cv2.getPerspectiveTransform(src, dst): 8 x 8 linear system, Gaussian elimination with partial pivoting (double)
__device__ bool perspective_from_quads(const float* src, const float* dst, double M[9]) {
  double A[8][9];
  for (int i = 0; i < 4; ++i) {
    const double x = src[2 * i], y = src[2 * i + 1], u = dst[2 * i], v = dst[2 * i + 1];
    double r0[9] = {x, y, 1, 0, 0, 0, -x * u, -y * u, u};
    double r1[9] = {0, 0, 0, x, y, 1, -x * v, -y * v, v};
    for (int j = 0; j < 9; ++j) { A[i][j] = r0[j]; A[i + 4][j] = r1[j]; }
  }
  for (int col = 0; col < 8; ++col) {
    int piv = col;
    for (int r = col + 1; r < 8; ++r) if (fabs(A[r][col]) > fabs(A[piv][col])) piv = r;
    if (fabs(A[piv][col]) < 1e-12) return false;
    if (piv != col) for (int j = 0; j < 9; ++j) { double t = A[col][j]; A[col][j] = A[piv][j]; A[piv][j] = t; }
    for (int r = 0; r < 8; ++r) {
      if (r == col) continue;
      const double f = A[r][col] / A[col][col];
      for (int j = col; j < 9; ++j) A[r][j] -= f * A[col][j];
    }
  }
  for (int i = 0; i < 8; ++i) M[i] = A[i][8] / A[i][i];
  M[8] = 1.0;
  return true;
}

// One block row per image: block (x, b).  out[b] = warpPerspective(img[b or 0], M, (out_w, out_h)), INTER_LINEAR, BORDER_CONSTANT 0.
__global__ void __launch_bounds__(256)
lp_unwarp_kernel(const unsigned char* __restrict__ img, int img_batch_stride, int H, int W, const float* __restrict__ corners, int out_h, int out_w,
                 unsigned char* __restrict__ out, int* __restrict__ ok) {
  const int b = blockIdx.y;
  __shared__ double Minv[9];
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    const float dstq[8] = {(float)out_w, (float)out_h, 0.f, (float)out_h, 0.f, 0.f, (float)out_w, 0.f};     // LP_corner, :390-393
    double M[9];
    // warpPerspective maps destination pixels through the INVERSE of M: solve the transform dst -> src directly
    s_ok = perspective_from_quads(dstq, corners + (size_t)b * 8, M) ? 1 : 0;
    for (int j = 0; j < 9; ++j) Minv[j] = M[j];
    if (ok && blockIdx.x == 0) ok[b] = s_ok;
  }
  __syncthreads();
  const unsigned char* src = img + (size_t)b * img_batch_stride;
  unsigned char* dst = out + (size_t)b * out_h * out_w * 3;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < out_h * out_w; p += gridDim.x * blockDim.x) {
    const int y = p / out_w, x = p - y * out_w;
    float v[3] = {0.f, 0.f, 0.f};
    if (s_ok) {
      const double w = Minv[6] * x + Minv[7] * y + Minv[8];
      const double sx = (Minv[0] * x + Minv[1] * y + Minv[2]) / w, sy = (Minv[3] * x + Minv[4] * y + Minv[5]) / w;
      const double fx0 = floor(sx), fy0 = floor(sy);
      const int x0 = (int)fx0, y0 = (int)fy0;
      const float ax = (float)(sx - fx0), ay = (float)(sy - fy0);
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int xi = x0 + dx, yi = y0 + dy;
          if ((unsigned)xi < (unsigned)W && (unsigned)yi < (unsigned)H) {
            const float wgt = (dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay);
            const unsigned char* q = src + ((size_t)yi * W + xi) * 3;
            v[0] += wgt * q[0]; v[1] += wgt * q[1]; v[2] += wgt * q[2];
          }
        }
    }
    for (int c = 0; c < 3; ++c) dst[(size_t)p * 3 + c] = (unsigned char)fminf(fmaxf(rintf(v[c]), 0.f), 255.f);
  }
}

// cv2.resize(img, (W, H)) for uint8 HWC frames, INTER_LINEAR - the exact fixed-point arithmetic of OpenCV's resize (coefficients in
// 1/2048, horizontal pass in int32, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2), so that the frame the network sees
// is bit-identical to what `cv2.resize` at car/video_node.py:150 produced (bit-exact when shrinking - the camera-frame case; +-1 on
// < 0.1 % of the values when enlarging).  One thread per output pixel (3 channels).
__global__ void __launch_bounds__(256)
resize_u8_kernel(const unsigned char* __restrict__ src, int sh, int sw, unsigned char* __restrict__ dst, int dh, int dw, int batch,
                 double scale_x, double scale_y) {
  const size_t total = (size_t)batch * dh * dw;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    const int dx = (int)(p % dw);
    const size_t t = p / dw;
    const int dy = (int)(t % dh), b = (int)(t / dh);
    float fx = (float)((dx + 0.5) * scale_x - 0.5), fy = (float)((dy + 0.5) * scale_y - 0.5);       // double, then float: OpenCV's order
    int sx = (int)floorf(fx), sy = (int)floorf(fy);
    fx -= (float)sx; fy -= (float)sy;
    if (sx < 0) { sx = 0; fx = 0.f; }
    if (sx >= sw - 1) { sx = sw - 1; fx = 0.f; }
    if (sy < 0) { sy = 0; fy = 0.f; }
    if (sy >= sh - 1) { sy = sh - 1; fy = 0.f; }
    const int a0 = __float2int_rn((1.f - fx) * 2048.f), a1 = __float2int_rn(fx * 2048.f);
    const int b0 = __float2int_rn((1.f - fy) * 2048.f), b1 = __float2int_rn(fy * 2048.f);
    const int sx1 = min(sx + 1, sw - 1), sy1 = min(sy + 1, sh - 1);
    const unsigned char* r0 = src + ((size_t)b * sh + sy) * sw * 3;
    const unsigned char* r1 = src + ((size_t)b * sh + sy1) * sw * 3;
    unsigned char* o = dst + p * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int S0 = r0[sx * 3 + c] * a0 + r0[sx1 * 3 + c] * a1;
      const int S1 = r1[sx * 3 + c] * a0 + r1[sx1 * 3 + c] * a1;
      o[c] = (unsigned char)((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2);
    }
  }
}

}  // namespace yb

using namespace yb;

extern "C" int yolo_resize_u8(const unsigned char* src, int batch, int src_h, int src_w, unsigned char* dst, int dst_h, int dst_w, void* stream) {
  if (!src || !dst || batch < 0 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return fail(YOLO_E_BADARG, "resize_u8: bad arguments");
  if (batch == 0) return YOLO_OK;
  const size_t total = (size_t)batch * dst_h * dst_w;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  // cv2: scale = 1 / (dsize / ssize) evaluated in double, applied in float
  const double inv_x = (double)dst_w / src_w, inv_y = (double)dst_h / src_h;
  resize_u8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, src_h, src_w, dst, dst_h, dst_w, batch, 1.0 / inv_x, 1.0 / inv_y);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

extern "C" int yolo_azimuth(const float* rows, int batch, int row_len, int n_class, float* out_angle, float* out_radius, void* stream) {
  if (!rows || !out_angle || batch < 0 || n_class < 1 || n_class > row_len || 360 % n_class) return fail(YOLO_E_BADARG, "azimuth: bad arguments");
  if (batch == 0) return YOLO_OK;
  azimuth_kernel<<<(batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(rows, batch, row_len, n_class, out_angle, out_radius);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

extern "C" int yolo_lp_corners(const float* poses, int batch, int pose_stride, int pose_offset, const double intrinsics[4], float x_scale,
                               float y_scale, float* out_corners, void* stream) {
  if (!poses || !intrinsics || !out_corners || batch < 0 || pose_stride < pose_offset + 6) return fail(YOLO_E_BADARG, "lp_corners: bad arguments");
  if (batch == 0) return YOLO_OK;
  lp_corners_kernel<<<(batch + 63) / 64, 64, 0, (cudaStream_t)stream>>>(poses, batch, pose_stride, pose_offset, intrinsics[0], intrinsics[1], intrinsics[2],
                                                                     intrinsics[3], x_scale, y_scale, out_corners);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}

extern "C" int yolo_lp_unwarp(const unsigned char* img, int batch, int img_is_batched, int h, int w, const float* corners, int out_h, int out_w,
                              unsigned char* out, int32_t* ok, void* stream) {
  if (!img || !corners || !out || batch < 0 || h < 1 || w < 1 || out_h < 1 || out_w < 1) return fail(YOLO_E_BADARG, "lp_unwarp: bad arguments");
  if (batch == 0) return YOLO_OK;
  const int bx = (out_h * out_w + 255) / 256 < 64 ? (out_h * out_w + 255) / 256 : 64;
  lp_unwarp_kernel<<<dim3(bx, batch), 256, 0, (cudaStream_t)stream>>>(img, img_is_batched ? h * w * 3 : 0, h, w, corners, out_h, out_w, out, ok);
  ++g_launches;
  YB_CUDA(cudaGetLastError());
  return YOLO_OK;
}
