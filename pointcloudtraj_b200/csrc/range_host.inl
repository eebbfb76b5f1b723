extern "C" int pc_range_batch(pc_index *ix, const float *q_xyz, int64_t m, int64_t q_stride, int space,
                              const double *range, int range_is_scalar,
                              int64_t *out_offsets, int32_t *out_idx, int64_t cap)
{
    (void)q_xyz; (void)m; (void)q_stride; (void)space; (void)range; (void)range_is_scalar; (void)out_offsets; (void)out_idx; (void)cap;
    return pc_fail(ix, PC_ENOTIMPL, "pc_range_batch: not implemented yet");
}
