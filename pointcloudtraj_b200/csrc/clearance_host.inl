extern "C" int pc_clearance_batch(pc_index *ix, const pc_traj *traj, int64_t n_traj,
                                  const int32_t *seg_order, const double *seg_T, const int64_t *seg_coef_off,
                                  int64_t n_seg, const double *coef, int64_t n_coef, int space,
                                  double dt, double horizon, const pc_radius_params *params,
                                  int32_t *out_first_hit, float *out_min_radius, int32_t *out_n_samples)
{
    (void)traj; (void)n_traj; (void)seg_order; (void)seg_T; (void)seg_coef_off; (void)n_seg; (void)coef; (void)n_coef; (void)space;
    (void)dt; (void)horizon; (void)params; (void)out_first_hit; (void)out_min_radius; (void)out_n_samples;
    return pc_fail(ix, PC_ENOTIMPL, "pc_clearance_batch: not implemented yet");
}
