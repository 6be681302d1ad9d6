// se3.cuh -- SE(3)/SO(3) math used on the device (FP64).
//
// Semantics follow the Sophus routines the reference calls (ref: include/third_party/sophus):
//   exp  : se3.hpp:761-783 + so3.hpp:583-619 (half-angle quaternion, Taylor branch below eps = 1e-10,
//          translation through the left Jacobian V)
//   log  : se3.hpp:223-257 + so3.hpp:247-290 (atan-based, V^-1)
//   mul  : so3.hpp:325-340 (+ quaternion re-normalisation of the SO3 constructor) and se3.hpp:304-312
//   ctor : Sophus::SE3d(Matrix4) -> Eigen Quaternion(Matrix3) (Shoemake branch on the trace)
// State is kept as (unit quaternion w,x,y,z ; translation) = 7 doubles ("Pose7").
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace tloam {

#define TL_HD __host__ __device__ __forceinline__

struct Pose7 {
  double qw, qx, qy, qz, tx, ty, tz;
};

struct Rt {          // rotation matrix (row-major) + translation, what the per-feature kernels consume
  double r[9];
  double t[3];
};

constexpr double kLieEps = 1e-10;  // sophus common.hpp:94

TL_HD void quat_to_rot(const Pose7& p, double r[9]) {
  const double x2 = p.qx + p.qx, y2 = p.qy + p.qy, z2 = p.qz + p.qz;
  const double wx = x2 * p.qw, wy = y2 * p.qw, wz = z2 * p.qw;
  const double xx = x2 * p.qx, xy = y2 * p.qx, xz = z2 * p.qx;
  const double yy = y2 * p.qy, yz = z2 * p.qy, zz = z2 * p.qz;
  r[0] = 1.0 - (yy + zz); r[1] = xy - wz;         r[2] = xz + wy;
  r[3] = xy + wz;         r[4] = 1.0 - (xx + zz); r[5] = yz - wx;
  r[6] = xz - wy;         r[7] = yz + wx;         r[8] = 1.0 - (xx + yy);
}

TL_HD Rt pose_to_rt(const Pose7& p) {
  Rt o;
  quat_to_rot(p, o.r);
  o.t[0] = p.tx; o.t[1] = p.ty; o.t[2] = p.tz;
  return o;
}

TL_HD void rt_apply(const Rt& T, double x, double y, double z, double& ox, double& oy, double& oz) {
  ox = T.r[0] * x + T.r[1] * y + T.r[2] * z + T.t[0];
  oy = T.r[3] * x + T.r[4] * y + T.r[5] * z + T.t[1];
  oz = T.r[6] * x + T.r[7] * y + T.r[8] * z + T.t[2];
}

// exp: tangent (upsilon, omega) -> pose
TL_HD Pose7 se3_exp(const double a[6]) {
  const double ox = a[3], oy = a[4], oz = a[5];
  const double th2 = ox * ox + oy * oy + oz * oz;
  Pose7 p;
  double theta, kim, kre;
  if (th2 < kLieEps * kLieEps) {
    theta = 0.0;
    const double th4 = th2 * th2;
    kim = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    kre = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    theta = sqrt(th2);
    double sh, ch;
    sincos(0.5 * theta, &sh, &ch);
    kim = sh / theta;
    kre = ch;
  }
  p.qw = kre; p.qx = kim * ox; p.qy = kim * oy; p.qz = kim * oz;
  // t = V * upsilon,  V = I + c1*W + c2*W^2 ;  W u = omega x u
  const double ux = a[0], uy = a[1], uz = a[2];
  if (theta < kLieEps) {
    double r[9];
    quat_to_rot(p, r);   // Sophus uses V = R in this branch
    p.tx = r[0] * ux + r[1] * uy + r[2] * uz;
    p.ty = r[3] * ux + r[4] * uy + r[5] * uz;
    p.tz = r[6] * ux + r[7] * uy + r[8] * uz;
  } else {
    // sin(theta), 1 - cos(theta) from the half-angle values already at hand (one sincos per exp)
    const double sh = kim * theta, ch = kre;
    const double s = 2.0 * sh * ch;
    const double c1 = 2.0 * sh * sh / th2;          // (1 - cos theta) / theta^2
    const double c2 = (theta - s) / (th2 * theta);
    const double wx = oy * uz - oz * uy, wy = oz * ux - ox * uz, wz = ox * uy - oy * ux;      // W u
    const double vx = oy * wz - oz * wy, vy = oz * wx - ox * wz, vz = ox * wy - oy * wx;      // W^2 u
    p.tx = ux + c1 * wx + c2 * vx;
    p.ty = uy + c1 * wy + c2 * vy;
    p.tz = uz + c1 * wz + c2 * vz;
  }
  return p;
}

// log: pose -> tangent
TL_HD void se3_log(const Pose7& p, double out[6]) {
  const double n2 = p.qx * p.qx + p.qy * p.qy + p.qz * p.qz;
  const double w = p.qw;
  double k, theta;  // omega = k * q.vec
  if (n2 < kLieEps * kLieEps) {
    k = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w);
    theta = 2.0 * n2 / w;
  } else {
    const double n = sqrt(n2);
    if (fabs(w) < kLieEps) {
      k = (w > 0.0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    } else {
      k = 2.0 * atan(n / w) / n;
    }
    theta = k * n;
  }
  const double ox = k * p.qx, oy = k * p.qy, oz = k * p.qz;
  // upsilon = V^-1 t ,  V^-1 = I - W/2 + c W^2
  double c;
  if (fabs(theta) < kLieEps) {
    c = 1.0 / 12.0;
  } else {
    // cos(theta/2) / sin(theta/2) == w / n for a unit quaternion with theta = 2 atan(n / w): no sincos needed
    c = (1.0 - 0.5 * theta * w / sqrt(n2)) / (theta * theta);
  }
  const double tx = p.tx, ty = p.ty, tz = p.tz;
  const double wx = oy * tz - oz * ty, wy = oz * tx - ox * tz, wz = ox * ty - oy * tx;
  const double vx = oy * wz - oz * wy, vy = oz * wx - ox * wz, vz = ox * wy - oy * wx;
  out[0] = tx - 0.5 * wx + c * vx;
  out[1] = ty - 0.5 * wy + c * vy;
  out[2] = tz - 0.5 * wz + c * vz;
  out[3] = ox; out[4] = oy; out[5] = oz;
}

// group product a * b (with the quaternion re-normalisation Sophus applies)
TL_HD Pose7 se3_mul(const Pose7& a, const Pose7& b) {
  Pose7 o;
  double w = a.qw * b.qw - a.qx * b.qx - a.qy * b.qy - a.qz * b.qz;
  double x = a.qw * b.qx + a.qx * b.qw + a.qy * b.qz - a.qz * b.qy;
  double y = a.qw * b.qy + a.qy * b.qw + a.qz * b.qx - a.qx * b.qz;
  double z = a.qw * b.qz + a.qz * b.qw + a.qx * b.qy - a.qy * b.qx;
  const double inv = 1.0 / sqrt(w * w + x * x + y * y + z * z);
  o.qw = w * inv; o.qx = x * inv; o.qy = y * inv; o.qz = z * inv;
  double r[9];
  quat_to_rot(a, r);
  o.tx = a.tx + r[0] * b.tx + r[1] * b.ty + r[2] * b.tz;
  o.ty = a.ty + r[3] * b.tx + r[4] * b.ty + r[5] * b.tz;
  o.tz = a.tz + r[6] * b.tx + r[7] * b.ty + r[8] * b.tz;
  return o;
}

// 4x4 column-major rigid matrix -> pose.  Returns false when the upper-left block is not a rotation
// (the reference would abort inside Sophus, so3.hpp:469-472).
TL_HD bool pose_from_matrix(const double T[16], Pose7& p) {
#define TLM(r, c) T[(c) * 4 + (r)]
  // orthogonality / handedness check
  double err = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += TLM(i, k) * TLM(j, k);
      const double e = fabs(s - (i == j ? 1.0 : 0.0));
      err = e > err ? e : err;
    }
  const double det = TLM(0, 0) * (TLM(1, 1) * TLM(2, 2) - TLM(1, 2) * TLM(2, 1)) -
                     TLM(0, 1) * (TLM(1, 0) * TLM(2, 2) - TLM(1, 2) * TLM(2, 0)) +
                     TLM(0, 2) * (TLM(1, 0) * TLM(2, 1) - TLM(1, 1) * TLM(2, 0));
  const bool ok = (err < 1e-9) && (det > 0.0) && isfinite(TLM(0, 3)) && isfinite(TLM(1, 3)) && isfinite(TLM(2, 3));
  double tr = TLM(0, 0) + TLM(1, 1) + TLM(2, 2);
  if (tr > 0.0) {
    double s = sqrt(tr + 1.0);
    p.qw = 0.5 * s;
    s = 0.5 / s;
    p.qx = (TLM(2, 1) - TLM(1, 2)) * s;
    p.qy = (TLM(0, 2) - TLM(2, 0)) * s;
    p.qz = (TLM(1, 0) - TLM(0, 1)) * s;
  } else {
    int i = 0;
    if (TLM(1, 1) > TLM(0, 0)) i = 1;
    if (TLM(2, 2) > TLM(i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(TLM(i, i) - TLM(j, j) - TLM(k, k) + 1.0);
    double v[3];
    v[i] = 0.5 * s;
    s = 0.5 / s;
    p.qw = (TLM(k, j) - TLM(j, k)) * s;
    v[j] = (TLM(j, i) + TLM(i, j)) * s;
    v[k] = (TLM(k, i) + TLM(i, k)) * s;
    p.qx = v[0]; p.qy = v[1]; p.qz = v[2];
  }
  p.tx = TLM(0, 3); p.ty = TLM(1, 3); p.tz = TLM(2, 3);
#undef TLM
  return ok;
}

TL_HD void pose_to_matrix(const Pose7& p, double T[16]) {
  double r[9];
  quat_to_rot(p, r);
  T[0] = r[0]; T[1] = r[3]; T[2] = r[6]; T[3] = 0.0;
  T[4] = r[1]; T[5] = r[4]; T[6] = r[7]; T[7] = 0.0;
  T[8] = r[2]; T[9] = r[5]; T[10] = r[8]; T[11] = 0.0;
  T[12] = p.tx; T[13] = p.ty; T[14] = p.tz; T[15] = 1.0;
}

}  // namespace tloam
