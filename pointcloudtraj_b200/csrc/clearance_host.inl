// clearance_host.inl -- host side of pc_clearance_batch (included by pc_index.cu).

static bool g_binom_uploaded[64] = { false };
#define PC_CLR_STAGE_BYTES ((int64_t)1 << 18)   // pinned staging block for small PC_HOST calls (one copy in, one out)
#define PC_CLR_FLAT_MAX ((int64_t)1 << 27)      // samples the flat path's schedule may hold (12 B each)

// C(n,k) exactly as Bernstein::setParam computes it (int factorial quotient, Planner/src/bezier_base.cpp:35-48)
static int pc_upload_binomials(pc_index *ix)
{
    if (ix->device < 64 && g_binom_uploaded[ix->device]) return PC_OK;
    double h[PC_MAX_ORDER + 1][PC_MAX_ORDER + 1];
    memset(h, 0, sizeof h);
    for (int n = 0; n <= PC_MAX_ORDER; n++) {
        for (int k = 0; k <= n; k++) {
            int fn = 1, fk = 1, fnk = 1;
            for (int i = n; i > 0; i--) fn *= i;
            for (int i = k; i > 0; i--) fk *= i;
            for (int i = n - k; i > 0; i--) fnk *= i;
            h[n][k] = (double)(fn / (fk * fnk));
        }
    }
    PC_CUDA(ix, cudaMemcpyToSymbol(pc_binom, h, sizeof h));
    if (ix->device < 64) g_binom_uploaded[ix->device] = true;
    return PC_OK;
}

extern "C" int pc_clearance_batch(pc_index *ix, const pc_traj *traj, int64_t n_traj,
                                  const int32_t *seg_order, const double *seg_T, const int64_t *seg_coef_off,
                                  int64_t n_seg, const double *coef, int64_t n_coef, int space,
                                  double dt, double horizon, const pc_radius_params *params,
                                  int32_t *out_first_hit, float *out_min_radius, int32_t *out_n_samples)
{
    if (!ix) return PC_EINVAL;
    if (n_traj < 0 || n_seg < 0 || n_coef < 0 || (space != PC_HOST && space != PC_DEVICE) ||
        (n_traj > 0 && !traj) || (n_seg > 0 && (!seg_order || !seg_T || !seg_coef_off || !coef)))
        return pc_fail(ix, PC_EINVAL, "pc_clearance_batch: bad argument");
    if (!(dt > 0.0) || !(horizon == horizon) || horizon / dt > 1e8)
        return pc_fail(ix, PC_EINVAL, "pc_clearance_batch: need dt > 0 and horizon / dt <= 1e8");
    if (n_traj > 0x7fffffff) return pc_fail(ix, PC_EINVAL, "pc_clearance_batch: at most 2^31-1 trajectories per call");
    pc_radius_dev R;
    int rc = pc_make_radius_dev(ix, params, PC_RADIUS_BOUNDED, &R);
    if (rc != PC_OK) return rc;
    PC_CUDA(ix, cudaSetDevice(ix->device));
    if ((rc = pc_upload_binomials(ix)) != PC_OK) return rc;
    if (n_traj == 0) return PC_OK;
    cudaStream_t st = ix->stream;

    if (space == PC_HOST) {
        // validate what the kernel will index with (host data only; device callers are trusted like any kernel argument)
        for (int64_t t = 0; t < n_traj; t++) {
            if (traj[t].num_seg < 0 || traj[t].first_seg < 0 || (int64_t)traj[t].first_seg + traj[t].num_seg > n_seg)
                return pc_fail(ix, PC_EINVAL, "pc_clearance_batch: trajectory %lld addresses segments outside [0, %lld)", (long long)t, (long long)n_seg);
        }
        for (int64_t s = 0; s < n_seg; s++) {
            if (seg_order[s] < 1 || seg_order[s] > PC_MAX_ORDER)
                return pc_fail(ix, PC_EINVAL, "pc_clearance_batch: segment %lld has order %d (supported 1..%d)", (long long)s, seg_order[s], PC_MAX_ORDER);
            if (seg_coef_off[s] < 0 || seg_coef_off[s] + 3 * (seg_order[s] + 1) > n_coef)
                return pc_fail(ix, PC_EINVAL, "pc_clearance_batch: segment %lld coefficients outside [0, %lld)", (long long)s, (long long)n_coef);
            if (!(seg_T[s] > 0.0)) return pc_fail(ix, PC_EINVAL, "pc_clearance_batch: segment %lld has T <= 0", (long long)s);
        }
    }

    const pc_traj_dev *d_traj = (const pc_traj_dev *)traj;
    const int32_t *d_order = seg_order; const double *d_T = seg_T; const int64_t *d_coff = seg_coef_off; const double *d_coef = coef;
    int32_t *d_fh = out_first_hit; float *d_mr = out_min_radius; int32_t *d_ns = out_n_samples;
    bool small = false;
    if (space == PC_HOST) {
        const int64_t b_traj = pc_align_up(n_traj * (int64_t)sizeof(pc_traj_dev), 256);
        const int64_t b_order = pc_align_up(n_seg * 4, 256), b_T = pc_align_up(n_seg * 8, 256), b_coff = pc_align_up(n_seg * 8, 256);
        const int64_t b_coef = pc_align_up(n_coef * 8, 256), b_out = pc_align_up(n_traj * 4, 256);
        void *base = nullptr;
        if ((rc = pc_scratch(ix, b_traj + b_order + b_T + b_coff + b_coef + 3 * b_out, &base)) != PC_OK) return rc;
        char *p = (char *)base;
        // the planner's own call (ONE trajectory, a few segments): the five arrays go through a pinned staging block in one copy
        // instead of five pageable ones (each of which costs ~10 us), the three result arrays come back the same way
        const int64_t up_bytes = b_traj + b_order + b_T + b_coff + b_coef;
        small = up_bytes <= PC_CLR_STAGE_BYTES && 3 * b_out <= PC_CLR_STAGE_BYTES;
        if (small && !ix->h_stage) PC_CUDA(ix, cudaHostAlloc((void **)&ix->h_stage, 2 * PC_CLR_STAGE_BYTES, cudaHostAllocDefault));
        char *hp = small ? ix->h_stage : nullptr;
#define PC_UP(dst, src, bytes, slot) do { dst = (decltype(dst))p; \
            if (small) { if ((bytes) > 0) memcpy(hp + (p - (char *)base), src, (size_t)(bytes)); } \
            else PC_CUDA(ix, cudaMemcpyAsync((void *)p, src, (size_t)(bytes), cudaMemcpyHostToDevice, st)); p += slot; } while (0)
        PC_UP(d_traj, traj, n_traj * sizeof(pc_traj_dev), b_traj);
        PC_UP(d_order, seg_order, n_seg * 4, b_order);
        PC_UP(d_T, seg_T, n_seg * 8, b_T);
        PC_UP(d_coff, seg_coef_off, n_seg * 8, b_coff);
        PC_UP(d_coef, coef, n_coef * 8, b_coef);
#undef PC_UP
        if (small) PC_CUDA(ix, cudaMemcpyAsync(base, hp, (size_t)up_bytes, cudaMemcpyHostToDevice, st));
        d_fh = (int32_t *)p; p += b_out;
        d_mr = (float *)p; p += b_out;
        d_ns = (int32_t *)p;
    }
    // The flat path (one warp per 32 samples of any trajectory) whenever its schedule fits PC_CLR_FLAT_MAX samples; a trajectory
    // never has more than horizon / dt + 1 samples (t_accu grows by dt per sample and stops at the horizon).
    const int64_t stride = horizon > 0.0 ? (int64_t)(horizon / dt) + 4 : 4;
    const int64_t ppt = (stride + 31) / 32;
    static const bool flat_on = []() { const char *v = getenv("PC_CLEARANCE_FLAT"); return !v || atoi(v) != 0; }();
    if (flat_on && n_traj * stride <= PC_CLR_FLAT_MAX) {
        const int64_t b_t = pc_align_up(n_traj * stride * 8, 256), b_s = pc_align_up(n_traj * stride * 4, 256);
        const int64_t b_pm = pc_align_up(n_traj * ppt * 8, 256), b_pf = pc_align_up(n_traj * ppt * 4, 256), b_n = pc_align_up(n_traj * 4, 256);
        pc_lane &L = ix->lane[0];
        // the schedule lives in the (idle) lane-0 float staging buffer: ix->scratch may hold the caller's uploaded arrays
        if ((rc = pc_grow(ix, &L.d_f32, &L.f32_cap, (b_t + b_s + b_pm + b_pf + b_n) / 4 + 64)) != PC_OK) return rc;
        char *p = (char *)L.d_f32;
        double *d_st = (double *)p; p += b_t;
        double *d_pm = (double *)p; p += b_pm;
        int32_t *d_ss = (int32_t *)p; p += b_s;
        int32_t *d_pf = (int32_t *)p; p += b_pf;
        int32_t *d_cnt = d_ns ? d_ns : (int32_t *)p;
        pc_clearance_schedule_kernel<<<(int)((n_traj + 63) / 64), 64, 0, st>>>(d_traj, n_traj, d_T, dt, horizon, stride, d_st, d_ss, d_cnt);
        const int64_t n_pk = n_traj * ppt;
        pc_clearance_eval_kernel<<<(int)((n_pk + PC_CLR_THREADS / 32 - 1) / (PC_CLR_THREADS / 32)), PC_CLR_THREADS, 0, st>>>(
            pc_tree_of(ix), R, n_traj, stride, ppt, d_order, d_T, d_coff, d_coef, d_st, d_ss, d_cnt, d_pm, d_pf);
        pc_clearance_finish_kernel<<<(int)((n_traj + 127) / 128), 128, 0, st>>>(n_traj, ppt, d_cnt, d_pm, d_pf, d_fh, d_mr);
        ix->launches += 3;
    } else {
        pc_clearance_kernel<<<(int)((n_traj + PC_CLR_THREADS / 32 - 1) / (PC_CLR_THREADS / 32)), PC_CLR_THREADS, 0, st>>>(pc_tree_of(ix), R, d_traj, n_traj, d_order, d_T, d_coff, d_coef,
                                                                   dt, horizon, d_fh, d_mr, d_ns);
        ix->launches++;
    }
    PC_CHECK_LAUNCH(ix);
    if (space == PC_HOST && small) {
        const int64_t b_out = pc_align_up(n_traj * 4, 256);
        char *hb = ix->h_stage + PC_CLR_STAGE_BYTES;
        PC_CUDA(ix, cudaMemcpyAsync(hb, d_fh, (size_t)(3 * b_out), cudaMemcpyDeviceToHost, st));      // first_hit | min_radius | n_samples
        PC_CUDA(ix, cudaStreamSynchronize(st));
        if (out_first_hit) memcpy(out_first_hit, hb, (size_t)n_traj * 4);
        if (out_min_radius) memcpy(out_min_radius, hb + b_out, (size_t)n_traj * 4);
        if (out_n_samples) memcpy(out_n_samples, hb + 2 * b_out, (size_t)n_traj * 4);
    } else if (space == PC_HOST) {
        if (out_first_hit) PC_CUDA(ix, cudaMemcpyAsync(out_first_hit, d_fh, (size_t)n_traj * 4, cudaMemcpyDeviceToHost, st));
        if (out_min_radius) PC_CUDA(ix, cudaMemcpyAsync(out_min_radius, d_mr, (size_t)n_traj * 4, cudaMemcpyDeviceToHost, st));
        if (out_n_samples) PC_CUDA(ix, cudaMemcpyAsync(out_n_samples, d_ns, (size_t)n_traj * 4, cudaMemcpyDeviceToHost, st));
        PC_CUDA(ix, cudaStreamSynchronize(st));
    }
    return PC_OK;
}
