// solver.cuh -- on-device trust-region state machine (single thread of the last block of k_eval).
//
// Restates, for ONE 6-parameter block, the control flow of Ceres Solver 2.0
//   TrustRegionMinimizer::Minimize / IterationZero / ComputeTrustRegionStep / HandleSuccessfulStep /
//   HandleInvalidStep / ParameterToleranceReached / FunctionToleranceReached (trust_region_minimizer.cc),
//   DoglegStrategy with SUBSPACE_DOGLEG (dogleg_strategy.cc), TrustRegionStepEvaluator (monotonic),
// with the options the reference sets (ref: registration.cpp:1036-1047) and Ceres defaults otherwise,
// followed by the reference's GNC bookkeeping (ref: registration.cpp:1049-1121).
//
// Everything Ceres derives from the Jacobian is derived here from the 6x6 normal equations
// H = J^T J, g = J^T r (robustified, unscaled) that k_eval reduces: with S = jacobi scaling,
//   J_s^T J_s = S H S,  J_s^T r = S g,  ||J_s col||^2 = (S H S)_jj,
//   model_cost_change = -(g_s . step + step^T H_s step / 2),
//   DENSE_QR on [J_s ; sqrt(mu) D] y = [r ; 0]  ==  (H_s + mu D^2) y = g_s  (solved by Cholesky, FP64).
#pragma once
#include "registration.cuh"

namespace tloam {

__device__ __forceinline__ double norm6(const double* v) {
  double s = 0.0;
  for (int i = 0; i < 6; ++i) s += v[i] * v[i];
  return sqrt(s);
}

// Cholesky solve of the SPD 6x6 system A y = b. Returns false if a pivot is not positive / finite.
__device__ __forceinline__ bool chol_solve6(const double A[36], const double b[6], double y[6]) {
  double L[36];
  for (int i = 0; i < 6; ++i) {
    for (int j = 0; j <= i; ++j) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      if (i == j) {
        if (!(s > 0.0) || !isfinite(s)) return false;
        L[i * 6 + i] = sqrt(s);
      } else {
        L[i * 6 + j] = s / L[j * 6 + j];
      }
    }
  }
  double z[6];
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * z[k];
    z[i] = s / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * y[k];
    y[i] = s / L[i * 6 + i];
  }
  for (int i = 0; i < 6; ++i) if (!isfinite(y[i])) return false;
  return true;
}

// minimise 0.5 y^T B y + g^T y on |y| = radius (2-D): the boundary problem of the subspace dogleg
// (dogleg_strategy.cc FindMinimumOnTrustRegionBoundary). Angular bracketing + bisection on f'.
__device__ __noinline__ void min_on_boundary_2d(const double B[4], const double g[2], double radius, double y[2]) {
  const double b01 = 0.5 * (B[1] + B[2]);
  double best_t = 0.0, best_f = DBL_MAX;
  const int N = 1024;
  const double two_pi = 6.283185307179586476925;
  auto f = [&](double t) {
    double s, c; sincos(t, &s, &c);
    c *= radius; s *= radius;
    return 0.5 * (B[0] * c * c + 2.0 * b01 * c * s + B[3] * s * s) + g[0] * c + g[1] * s;
  };
  auto df = [&](double t) {
    double s, c; sincos(t, &s, &c);
    return radius * radius * ((B[3] - B[0]) * c * s + b01 * (c * c - s * s)) + radius * (-g[0] * s + g[1] * c);
  };
  double d_prev = df(0.0), f_prev = f(0.0);
  for (int i = 0; i < N; ++i) {
    const double t0 = two_pi * i / N, t1 = two_pi * (i + 1) / N;
    const double d1 = df(t1), f1 = f(t1);
    double cand_t, cand_f;
    if (d_prev < 0.0 && d1 > 0.0) {
      double lo = t0, hi = t1;
      for (int it = 0; it < 100; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (df(mid) < 0.0) lo = mid; else hi = mid;
      }
      cand_t = 0.5 * (lo + hi); cand_f = f(cand_t);
    } else if (f_prev < f1) { cand_t = t0; cand_f = f_prev; }
    else { cand_t = t1; cand_f = f1; }
    if (cand_f < best_f) { best_f = cand_f; best_t = cand_t; }
    d_prev = d1; f_prev = f1;
  }
  double s, c; sincos(best_t, &s, &c);
  y[0] = radius * c; y[1] = radius * s;
}

struct SolverIO {
  const DeviceCtx* ctx;
  FrameState* st;
  tloam_b200_stats* tr;
};

__device__ __forceinline__ tloam_b200_outer_trace* outer_trace(const SolverIO& io) {
  return (io.st->outer < TLOAM_B200_MAX_OUTER) ? &io.tr->outer[io.st->outer] : nullptr;
}
__device__ __forceinline__ tloam_b200_inner_trace* inner_trace(const SolverIO& io) {
  tloam_b200_outer_trace* ot = outer_trace(io);
  const int it = io.st->iter;
  return (ot && it >= 1 && it <= TLOAM_B200_MAX_INNER) ? &ot->inner[it - 1] : nullptr;
}

// |x - Plus(x, -g)|_inf  (TrustRegionMinimizer::EvaluateGradientAndJacobian)
__device__ __forceinline__ double gradient_max_norm(const FrameState* st, const double g[6]) {
  double ng[6], proj[6];
  for (int i = 0; i < 6; ++i) ng[i] = -g[i];
  se3_log(se3_mul(se3_exp(ng), st->xq), proj);
  double m = 0.0;
  for (int i = 0; i < 6; ++i) m = fmax(m, fabs(st->x[i] - proj[i]));
  return m;
}

__device__ void finish_frame(const SolverIO& io) {
  FrameState* st = io.st;
  const Pose7 fin = se3_exp(st->x);                     // ref: registration.cpp:1124
  pose_to_matrix(fin, st->result);
  for (int i = 0; i < 16; ++i) st->curr_pose[i] = st->result[i];
  for (int i = 0; i < 6; ++i) io.tr->x_final[i] = st->x[i];
  st->frame_done = 1;
}

// The Ceres solve of this outer iteration is over: GNC bookkeeping, ref: registration.cpp:1049-1121.
__device__ void end_of_solve(const SolverIO& io, int termination) {
  FrameState* st = io.st;
  const DeviceCtx& c = *io.ctx;
  tloam_b200_outer_trace* ot = outer_trace(io);
  const double mu = st->mu;
  st->th1 = (mu + 1.0) / mu * st->c2;                   // :1049
  st->th2 = mu / (mu + 1.0) * st->c2;                   // :1050
  st->mu_used = mu;
  if (ot) {
    ot->termination = termination;
    ot->final_cost = st->x_cost;
    ot->n_inner = st->iter;
    for (int i = 0; i < 6; ++i) ot->x_end[i] = st->x[i];
    ot->mu = mu; ot->th1 = st->th1; ot->th2 = st->th2;
    for (int k = 0; k < 4; ++k) ot->slot_sum[k] = st->slot_sum[k];
  }
  st->mu = mu * exp((double)(st->outer + 1) * c.gnc_factor);   // :1089
  io.tr->n_outer = st->outer + 1;
  const double planar_cost = st->slot_sum[kPlanar];            // :1094
  const double diff = fabs(planar_cost - st->planar_prev);     // :1096
  if (diff < c.cost_threshold) {                               // :1108
    io.tr->converged_early = 1;
    finish_frame(io);
    return;
  }
  st->planar_prev = planar_cost;                               // :1113
  st->outer += 1;
  if (st->outer >= c.max_iterations) { finish_frame(io); return; }
  st->phase = kPhaseIter0;        // next: k_correspond (weights updated + slots zeroed there), then k_eval<first>
  st->evalq = st->xq;
}

// DoglegStrategy::ComputeStep (first call after an accepted / invalid step): diagonal, scaled gradient,
// Gauss-Newton step, subspace model.  Returns false on LINEAR_SOLVER_FAILURE.
__device__ __noinline__ bool compute_model(FrameState* st) {
  double Hs[36], gs[6];
  for (int i = 0; i < 6; ++i) {
    gs[i] = st->scale[i] * st->g[i];
    for (int j = 0; j < 6; ++j) {
      const double h = (i <= j) ? st->H[tri(i, j)] : st->H[tri(j, i)];
      Hs[i * 6 + j] = st->scale[i] * h * st->scale[j];
    }
  }
  for (int i = 0; i < 6; ++i) {
    st->D[i] = sqrt(fmin(fmax(Hs[i * 6 + i], 1e-6), 1e32));    // min/max_lm_diagonal
    st->sgrad[i] = gs[i] / st->D[i];
  }
  bool ok = false;
  double y[6];
  while (st->mu_lm < 1.0) {                                     // kMaxMu
    double A[36];
    for (int i = 0; i < 36; ++i) A[i] = Hs[i];
    for (int i = 0; i < 6; ++i) A[i * 6 + i] += st->mu_lm * st->D[i] * st->D[i];
    if (chol_solve6(A, gs, y)) { ok = true; break; }
    st->mu_lm *= 10.0;                                          // mu_increase_factor_
  }
  if (!ok) return false;
  for (int i = 0; i < 6; ++i) st->gn[i] = -st->D[i] * y[i];
  st->gn_norm = norm6(st->gn);
  // subspace model (ComputeSubspaceModel): orthonormal basis of span{sgrad, gn}
  const double n0 = norm6(st->sgrad), n1 = st->gn_norm;
  const double* first = (n0 >= n1) ? st->sgrad : st->gn;
  const double* second = (n0 >= n1) ? st->gn : st->sgrad;
  const double nf = fmax(n0, n1);
  if (nf == 0.0) return false;
  double u0[6], u1[6], pr = 0.0;
  for (int i = 0; i < 6; ++i) { u0[i] = first[i] / nf; pr += u0[i] * second[i]; }
  for (int i = 0; i < 6; ++i) u1[i] = second[i] - pr * u0[i];
  const double n2 = norm6(u1);
  st->sub_1d = !(n2 > nf * 6.0 * DBL_EPSILON);
  if (!st->sub_1d) {
    double t0[6], t1[6];
    st->sub_g[0] = st->sub_g[1] = 0.0;
    for (int i = 0; i < 6; ++i) {
      u1[i] /= n2;
      st->sub_basis[i] = u0[i]; st->sub_basis[6 + i] = u1[i];
      st->sub_g[0] += u0[i] * st->sgrad[i]; st->sub_g[1] += u1[i] * st->sgrad[i];
      t0[i] = u0[i] / st->D[i]; t1[i] = u1[i] / st->D[i];
    }
    double b00 = 0, b01 = 0, b11 = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        b00 += t0[i] * Hs[i * 6 + j] * t0[j];
        b01 += t0[i] * Hs[i * 6 + j] * t1[j];
        b11 += t1[i] * Hs[i * 6 + j] * t1[j];
      }
    st->sub_B[0] = b00; st->sub_B[1] = st->sub_B[2] = b01; st->sub_B[3] = b11;
  }
  return true;
}

// Loop "compute step -> candidate" until a candidate needs a fresh evaluation (returns with
// phase = kPhaseCand) or the solve terminates (end_of_solve called).
__device__ __noinline__ void advance(const SolverIO& io) {
  FrameState* st = io.st;
  const DeviceCtx& c = *io.ctx;
  while (true) {
    if (st->iter >= c.ceres_max_it) { end_of_solve(io, 0); return; }   // MaxSolverIterationsReached
    st->iter += 1;
    tloam_b200_inner_trace* it = inner_trace(io);
    bool solver_ok = true;
    if (!st->reuse) {
      st->reuse = 1;
      solver_ok = compute_model(st);
    }
    double step[6] = {0, 0, 0, 0, 0, 0};
    if (solver_ok) {                                                   // ComputeSubspaceDoglegStep
      if (st->gn_norm <= st->radius) {
        for (int i = 0; i < 6; ++i) step[i] = st->gn[i] / st->D[i];
        st->step_norm = st->gn_norm; st->used_gn = 1;
      } else if (st->sub_1d) {
        const double gnm = norm6(st->sgrad);
        for (int i = 0; i < 6; ++i) step[i] = -(st->radius / gnm) * st->sgrad[i] / st->D[i];
        st->step_norm = st->radius; st->used_gn = 0;
      } else {
        double y2[2];
        min_on_boundary_2d(st->sub_B, st->sub_g, st->radius, y2);
        for (int i = 0; i < 6; ++i) step[i] = (st->sub_basis[i] * y2[0] + st->sub_basis[6 + i] * y2[1]) / st->D[i];
        st->step_norm = st->radius; st->used_gn = 0;
      }
    }
    // model cost change = -(g_s.step + step^T H_s step / 2)
    double mcc = 0.0;
    bool valid = false;
    if (solver_ok) {
      double lin = 0.0, quad = 0.0;
      for (int i = 0; i < 6; ++i) {
        lin += st->scale[i] * st->g[i] * step[i];
        for (int j = 0; j < 6; ++j) {
          const double h = (i <= j) ? st->H[tri(i, j)] : st->H[tri(j, i)];
          quad += step[i] * st->scale[i] * h * st->scale[j] * step[j];
        }
      }
      mcc = -(lin + 0.5 * quad);
      valid = mcc > 0.0;
    }
    st->model_cost_change = mcc;
    if (it) {
      it->radius = st->radius; it->step_norm_scaled = st->step_norm; it->used_gauss_newton = st->used_gn;
      it->model_cost_change = mcc; it->accepted = 0; it->relative_decrease = 0.0; it->candidate_cost = 0.0;
    }
    if (!valid) {                                                      // HandleInvalidStep
      st->num_invalid += 1;
      if (it) it->accepted = -1;
      if (st->num_invalid >= 5) { end_of_solve(io, 6); return; }
      st->mu_lm *= 10.0; st->reuse = 0;                                // StepIsInvalid
      continue;
    }
    st->num_invalid = 0;
    double delta[6];
    for (int i = 0; i < 6; ++i) delta[i] = step[i] * st->scale[i];     // undo the Jacobi scaling
    st->candq = se3_mul(se3_exp(delta), st->xq);                       // Plus: exp(delta) * exp(x)
    se3_log(st->candq, st->cand);
    if (it) for (int i = 0; i < 6; ++i) it->x_candidate[i] = st->cand[i];
    bool same = st->last_cand_valid != 0;
    for (int i = 0; i < 6 && same; ++i) same = (st->cand[i] == st->last_cand[i]);
    if (same) {
      // identical to the candidate that was just evaluated and rejected (reuse_ = true and the
      // Gauss-Newton step still fits the halved radius): the evaluation, both tolerance tests and the step
      // quality repeat exactly, so the step is rejected again without re-running the kernel.
      if (it) {
        it->candidate_cost = st->last_cand_cost;
        it->relative_decrease = (st->x_cost - st->last_cand_cost) / mcc;
        it->accepted = 0;
      }
      st->radius *= 0.5; st->reuse = 1;                                // StepRejected
      if (st->radius <= 1e-32) { end_of_solve(io, 4); return; }
      continue;
    }
    st->evalq = st->candq;
    st->phase = kPhaseCand;
    return;
  }
}

// Called by one thread after the per-block partials have been summed into tot[kNRed].
__device__ __noinline__ void solver_on_eval(const DeviceCtx& ctx, const double* tot) {
  SolverIO io{&ctx, ctx.st, ctx.stats};
  FrameState* st = io.st;
  const double cost = tot[27];
  for (int k = 0; k < 4; ++k) st->slot_sum[k] = tot[28 + k];
  tloam_b200_outer_trace* ot = outer_trace(io);
  if (!isfinite(cost)) { st->status = TLOAM_B200_ERR_NUMERIC; finish_frame(io); return; }

  if (st->phase == kPhaseIter0) {
    // ---- IterationZero ----
    st->iter = 0; st->num_invalid = 0; st->reuse = 0; st->last_cand_valid = 0;
    st->radius = 1e4; st->mu_lm = 1e-8;                                 // initial_trust_region_radius, kMinMu
    int nf_total = 0;
    for (int k = 0; k < 4; ++k) {
      const int nf = (int)(tot[32 + k] + 0.5);
      nf_total += nf;
      if (ot) ot->n_factors[k] = nf;
    }
    if (ot) for (int i = 0; i < 6; ++i) ot->x_start[i] = st->x[i];
    st->x_cost = cost;
    for (int i = 0; i < 21; ++i) st->H[i] = tot[i];
    for (int i = 0; i < 6; ++i) st->g[i] = tot[21 + i];
    if (ot) {
      ot->initial_cost = cost;
      for (int i = 0; i < 6; ++i) {
        ot->g0[i] = st->g[i];
        for (int j = 0; j < 6; ++j) ot->H0[i * 6 + j] = (i <= j) ? st->H[tri(i, j)] : st->H[tri(j, i)];
      }
    }
    if (nf_total == 0) { end_of_solve(io, 5); return; }                  // no residual blocks
    for (int i = 0; i < 6; ++i) st->scale[i] = 1.0 / (1.0 + sqrt(st->H[tri(i, i)]));   // jacobi scaling
    st->x_norm = norm6(st->x);
    if (gradient_max_norm(st, st->g) <= 1e-10) { end_of_solve(io, 3); return; }
    advance(io);
    return;
  }

  // ---- the evaluation was at the candidate ----
  tloam_b200_inner_trace* it = inner_trace(io);
  if (it) it->candidate_cost = cost;
  {
    double d[6];
    for (int i = 0; i < 6; ++i) d[i] = st->x[i] - st->cand[i];
    if (norm6(d) <= 1e-8 * (st->x_norm + 1e-8)) {                      // ParameterToleranceReached
      if (it) it->accepted = 2;
      end_of_solve(io, 2); return;
    }
  }
  if (fabs(st->x_cost - cost) <= 1e-6 * st->x_cost) {                  // FunctionToleranceReached
    if (it) it->accepted = 2;
    end_of_solve(io, 1); return;
  }
  const double rel = (st->x_cost - cost) / st->model_cost_change;      // StepQuality (monotonic)
  if (it) it->relative_decrease = rel;
  if (rel > 1e-3) {                                                    // min_relative_decrease
    // HandleSuccessfulStep
    for (int i = 0; i < 6; ++i) st->x[i] = st->cand[i];
    st->xq = st->candq;
    st->x_norm = norm6(st->x);
    st->x_cost = cost;
    for (int i = 0; i < 21; ++i) st->H[i] = tot[i];
    for (int i = 0; i < 6; ++i) st->g[i] = tot[21 + i];
    if (rel < 0.25) st->radius *= 0.5;                                 // DoglegStrategy::StepAccepted
    if (rel > 0.75) st->radius = fmax(st->radius, 3.0 * st->step_norm);
    st->mu_lm = fmax(1e-8, 2.0 * st->mu_lm / 10.0);
    st->reuse = 0; st->last_cand_valid = 0;
    if (it) it->accepted = 1;
    if (gradient_max_norm(st, st->g) <= 1e-10) { end_of_solve(io, 3); return; }
  } else {
    st->radius *= 0.5; st->reuse = 1;                                  // StepRejected
    for (int i = 0; i < 6; ++i) st->last_cand[i] = st->cand[i];
    st->last_cand_cost = cost; st->last_cand_valid = 1;
    if (it) it->accepted = 0;
  }
  if (st->radius <= 1e-32) { end_of_solve(io, 4); return; }           // MinTrustRegionRadiusReached
  advance(io);
}

}  // namespace tloam
