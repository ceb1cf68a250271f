// prb_convert.h — host side: compiled fp64 model (include/prb_model.h) -> fp32 DevModel.
#pragma once
#include <math.h>
#include <string.h>

#include <string>

#include "../../include/prb_model.h"
#include "prb_device.h"

static inline void prb_mat_to_quat_host(const double* R, float* q) {
  double t = R[0] + R[4] + R[8];
  double qq[4];
  if (t > 0) {
    double s = sqrt(t + 1.0);
    qq[3] = s * 0.5; s = 0.5 / s;
    qq[0] = (R[7] - R[5]) * s; qq[1] = (R[2] - R[6]) * s; qq[2] = (R[3] - R[1]) * s;
  } else {
    int i = R[0] < R[4] ? (R[4] < R[8] ? 2 : 1) : (R[0] < R[8] ? 2 : 0);
    int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    qq[i] = s * 0.5; s = 0.5 / s;
    qq[3] = (R[3 * k + j] - R[3 * j + k]) * s;
    qq[j] = (R[3 * j + i] + R[3 * i + j]) * s;
    qq[k] = (R[3 * k + i] + R[3 * i + k]) * s;
  }
  for (int i = 0; i < 4; i++) q[i] = (float)qq[i];
}

// returns empty string on success, else the reason the model cannot be represented
static inline std::string prb_convert_model(const prb_model* m, DevModel* D) {
  memset(D, 0, sizeof(*D));
  if (m->nd < 1 || m->nd > PRB_MAXD) return "nd out of range";
  if (m->n_free > PRB_MAXFREE || m->n_slide > PRB_MAXSLIDE) return "too many bodies";
  if (m->n_col > PRB_MAXCOL) return "too many colliders";
  if (m->n_pair > PRB_MAXPAIR) return "too many collider pairs";
  if (m->n_ik != 6 && m->n_ik != 7) return "n_ik must be 6 or 7";
  if (m->n_grip > 8 || m->goal_dim > 12 || m->obs_dim > 24 || m->fps_dim > 24 || m->observation_dim > 24) return "dims out of range";
  D->env_kind = m->env_kind; D->arm_kind = m->arm_kind; D->nd = m->nd; D->n_ik = m->n_ik;
  D->n_free = m->n_free; D->n_slide = m->n_slide; D->n_col = m->n_col; D->n_pair = m->n_pair; D->n_grip = m->n_grip;
  D->ik_calls = m->ik_calls; D->ik_iters = m->ik_iters; D->ik_reset_iters = m->ik_reset_iters;
  D->n_substeps = m->n_substeps; D->solver_iters = m->solver_iters; D->settle_steps = m->settle_steps;
  D->obs_dim = m->obs_dim; D->goal_dim = m->goal_dim; D->fps_dim = m->fps_dim; D->observation_dim = m->observation_dim;
  D->use_orientation = m->use_orientation; D->return_velocity = m->return_velocity; D->play = m->play;
  D->grip_obs_dof = m->grip_obs_dof; D->gear_a = m->gear_a; D->gear_b = m->gear_b;
  D->nv = m->nd + 6 * m->n_free + m->n_slide;
  if (D->nv > 32) return "more than 32 velocity DoF";
  D->state_dim = 5 * m->nd + 13 * m->n_free + 2 * m->n_slide + m->goal_dim + 10;
  D->state_stride = (D->state_dim + 31) / 32 * 32;
  const int nd = m->nd;
  for (int i = 0; i < nd; i++) {
    D->parent[i] = m->arm_parent[i]; D->jtype[i] = m->arm_jtype[i];
    if (D->parent[i] >= i) return "links must be topologically ordered";
    // path root -> i
    int chain[PRB_MAXDEPTH * 2], n = 0;
    for (int l = i; l >= 0; l = m->arm_parent[l]) { if (n >= PRB_MAXDEPTH) return "kinematic tree too deep"; chain[n++] = l; }
    D->depth[i] = n;
    for (int k = 0; k < n; k++) D->path[i][k] = (unsigned char)chain[n - 1 - k];
    unsigned anc = 0;
    for (int k = 0; k < n; k++) anc |= 1u << chain[k];
    D->anc_mask[i] = anc;
    for (int k = 0; k < 3; k++) { D->jpos[i][k] = (float)m->arm_jpos[3 * i + k]; D->axis[i][k] = (float)m->arm_axis[3 * i + k]; D->com[i][k] = (float)m->arm_com[3 * i + k]; }
    for (int k = 0; k < 9; k++) D->jrot[i][k] = (float)m->arm_jrot[9 * i + k];
    const double* I = m->arm_inertia + 9 * i;
    D->inertia[i][0] = (float)I[0]; D->inertia[i][1] = (float)I[1]; D->inertia[i][2] = (float)I[2];
    D->inertia[i][3] = (float)I[4]; D->inertia[i][4] = (float)I[5]; D->inertia[i][5] = (float)I[8];
    D->mass[i] = (float)m->arm_mass[i]; D->lo[i] = (float)m->arm_lo[i]; D->hi[i] = (float)m->arm_hi[i];
    D->jdamp[i] = (float)m->arm_jdamp[i]; D->rest[i] = (float)m->arm_rest[i];
  }
  for (int j = 0; j < nd; j++) {
    unsigned sub = 0;
    for (int i = 0; i < nd; i++) if ((D->anc_mask[i] >> j) & 1u) sub |= 1u << i;
    D->sub_mask[j] = sub;
  }
  // the IK kernels assume the first n_ik joints form a serial revolute chain carrying the EE site
  for (int i = 0; i < m->n_ik; i++) if (m->arm_parent[i] != i - 1 || m->arm_jtype[i] != 0) return "first n_ik joints must be a serial revolute chain";
  if (m->site_link[0] != m->n_ik - 1) return "end-effector site must sit on the last chain link";
  for (int k = 0; k < 3; k++) D->base_pos[k] = (float)m->arm_base_pos[k];
  for (int k = 0; k < 9; k++) D->base_rot[k] = (float)m->arm_base_rot[k];
  prb_mat_to_quat_host(m->arm_base_rot, D->base_quat);
  for (int s = 0; s < 4; s++) {
    D->site_link[s] = m->site_link[s];
    for (int k = 0; k < 3; k++) D->site_pos[s][k] = (float)m->site_pos[3 * s + k];
    for (int k = 0; k < 9; k++) D->site_rot[s][k] = (float)m->site_rot[9 * s + k];
  }
  for (int j = 0; j < 8; j++) D->joints_obs_dof[j] = m->joints_obs_dof[j];
  for (int c = 0; c < m->n_col; c++) {
    D->col_body[c] = (signed char)m->col_body[c]; D->col_link[c] = (signed char)m->col_link[c]; D->col_urdf[c] = (signed char)m->col_urdf_link[c];
    if (m->col_obj[c] < 0 || m->col_obj[c] > 255) return "collision object id out of range";
    D->col_obj[c] = (unsigned char)m->col_obj[c];
    for (int k = 0; k < 3; k++) { D->col_pos[c][k] = (float)m->col_pos[3 * c + k]; D->col_half[c][k] = (float)m->col_half[3 * c + k]; }
    for (int k = 0; k < 9; k++) D->col_rot[c][k] = (float)m->col_rot[9 * c + k];
    D->col_fric[c] = (float)m->col_friction[c]; D->col_spin[c] = (float)m->col_spin[c];
    D->col_stiff[c] = (float)m->col_stiffness[c]; D->col_damp[c] = (float)m->col_damping[c];
  }
  for (int k = 0; k < m->n_pair; k++) { D->pair_a[k] = (unsigned char)m->pair_a[k]; D->pair_b[k] = (unsigned char)m->pair_b[k]; }
  for (int b = 0; b < m->n_free; b++) {
    D->free_mass[b] = (float)m->free_mass[b]; D->free_ld[b] = (float)m->free_lin_damp[b]; D->free_ad[b] = (float)m->free_ang_damp[b];
    for (int k = 0; k < 3; k++) { D->free_inertia[b][k] = (float)m->free_inertia[3 * b + k]; D->free_pos0[b][k] = (float)m->free_pos0[3 * b + k]; }
    for (int k = 0; k < 4; k++) D->free_quat0[b][k] = (float)m->free_quat0[4 * b + k];
  }
  for (int s = 0; s < m->n_slide; s++) {
    D->slide_jtype[s] = m->slide_jtype[s];
    for (int k = 0; k < 3; k++) { D->slide_pos[s][k] = (float)m->slide_pos[3 * s + k]; D->slide_axis[s][k] = (float)m->slide_axis[3 * s + k]; }
    for (int k = 0; k < 9; k++) D->slide_rot[s][k] = (float)m->slide_rot[9 * s + k];
    for (int i = 0; i < 3; i++) {
      double a = 0;
      for (int k = 0; k < 3; k++) a += m->slide_rot[9 * s + 3 * i + k] * m->slide_axis[3 * s + k];
      D->slide_axis_w[s][i] = (float)a;
    }
    D->slide_minv[s] = (float)(m->slide_jtype[s] == 1 ? 1.0 / m->slide_mass[s] : 1.0 / m->slide_inertia[s]);
    D->slide_ad[s] = (float)m->slide_ang_damp[s];
    for (int k = 0; k < 4; k++) D->slide_motor[s][k] = (float)m->slide_motor[4 * s + k];
  }
  for (int i = 0; i < m->n_ik; i++) { D->ctrl_ll[i] = (float)m->ctrl_ll[i]; D->ctrl_ul[i] = (float)m->ctrl_ul[i]; D->ctrl_inc[i] = (float)m->ctrl_inc[i]; }
  for (int k = 0; k < m->n_grip; k++) {
    D->grip_dof[k] = m->grip_dof[k]; D->grip_mimic[k] = m->grip_mimic[k];
    D->grip_scale[k] = (float)m->grip_scale[k]; D->grip_offset[k] = (float)m->grip_offset[k]; D->grip_force[k] = (float)m->grip_force[k];
  }
  for (int k = 0; k < 3; k++) {
    D->goal_lo[k] = (float)m->goal_lo[k]; D->goal_hi[k] = (float)m->goal_hi[k]; D->obj_lo[k] = (float)m->obj_lo[k];
    D->obj_hi[k] = (float)m->obj_hi[k]; D->env_hi[k] = (float)m->env_hi[k];
  }
  for (int k = 0; k < 4; k++) D->default_orn[k] = (float)m->default_orn[k];
  for (int k = 0; k < PRB_NPARAM; k++) D->params[k] = (float)m->params[k];
  return "";
}
