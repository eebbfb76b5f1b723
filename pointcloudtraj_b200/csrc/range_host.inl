// range_host.inl -- host side of pc_range_batch (included by pc_index.cu).

static int pc_scratch(pc_index *ix, int64_t bytes, void **out)
{
    if (bytes > ix->scratch_cap) {
        // (see pc_grow: reallocations are what must stay rare, not bytes -- twice the request, twice the old size, 8 MB at least)
        int64_t c = 2 * bytes;
        if (c < 2 * ix->scratch_cap) c = 2 * ix->scratch_cap;
        if (c < (8 << 20)) c = 8 << 20;
        if (ix->scratch) { pc_pool_free(ix, ix->scratch); ix->scratch = nullptr; ix->scratch_cap = 0; }
        PC_CUDA(ix, pc_pool_alloc(ix, &ix->scratch, (size_t)c));
        ix->scratch_cap = c;
    }
    *out = ix->scratch;
    return PC_OK;
}

static inline int64_t pc_align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

extern "C" int pc_range_batch(pc_index *ix, const float *q_xyz, int64_t m, int64_t q_stride, int space,
                              const double *range, int range_is_scalar,
                              int64_t *out_offsets, int32_t *out_idx, int64_t cap)
{
    int rc = pc_check_query_args(ix, "pc_range_batch", q_xyz, m, q_stride, space);
    if (rc != PC_OK) return rc;
    if (space != PC_HOST && space != PC_DEVICE)            // the ASYNC spaces are for nearest / radius batches only
        return pc_fail(ix, PC_EINVAL, "pc_range_batch: space must be PC_HOST or PC_DEVICE");
    if (!out_offsets || (m > 0 && !range) || cap < 0 || (cap > 0 && !out_idx))
        return pc_fail(ix, PC_EINVAL, "pc_range_batch: bad argument");
    PC_CUDA(ix, cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    // PC_TRACE_SLOW_MS=<ms>: a call that takes longer reports where the host spent the time (stderr)
    struct pc_trace {
        double t[8]; int n = 0; double limit;
        static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
        void mark() { if (limit > 0.0 && n < 8) t[n++] = now(); }
    } tr;
    tr.limit = ix->trace_slow_ms;
    tr.mark();
    const int qs = (int)q_stride;
    const int64_t n_range = range_is_scalar ? 1 : m;
    const int64_t n_tiles = (m + PC_SCAN_TILE - 1) / PC_SCAN_TILE;

    // device scratch: [queries | ranges | offsets] (PC_HOST only) + counts + tile sums
    const int64_t b_q = space == PC_HOST ? pc_align_up(m * qs * (int64_t)sizeof(float), 256) : 0;
    const int64_t b_r = space == PC_HOST ? pc_align_up(n_range * (int64_t)sizeof(double), 256) : 0;
    const int64_t b_off = space == PC_HOST ? pc_align_up((m + 1) * (int64_t)sizeof(int64_t), 256) : 0;
    const int64_t b_cnt = pc_align_up((m + 1) * (int64_t)sizeof(int64_t), 256);
    const int64_t b_tile = pc_align_up((n_tiles + 1) * (int64_t)sizeof(int64_t), 256);
    const int64_t b_long = pc_align_up((m + 1) * (int64_t)sizeof(int64_t), 256);      // queue of lists too long for the in-warp sort
    // the counting pass of a call that wants the lists parks the short ones (<= PC_RCAP_HITS hits) in a staging buffer
    const bool capture = out_idx && cap > 0 && m > 0;
    const int64_t stage_cap = capture ? (cap < m * PC_RCAP_HITS ? cap : m * PC_RCAP_HITS) : 0;
    const int64_t b_stage = capture ? pc_align_up(stage_cap * (int64_t)sizeof(int32_t), 256) : 0;
    const int64_t b_pos = capture ? pc_align_up(m * (int64_t)sizeof(int64_t), 256) : 0;
    const int64_t b_todo = capture ? pc_align_up((m + 1) * (int64_t)sizeof(int64_t), 256) : 0;
    void *base = nullptr;
    if ((rc = pc_scratch(ix, b_q + b_r + b_off + b_cnt + b_tile + b_long + b_stage + b_pos + b_todo + 256, &base)) != PC_OK) return rc;
    tr.mark();            // 1: scratch
    char *p = (char *)base;
    const float *d_q = q_xyz; const double *d_r = range; int64_t *d_off = out_offsets;
    if (space == PC_HOST) {
        d_q = (const float *)p; p += b_q;
        d_r = (const double *)p; p += b_r;
        d_off = (int64_t *)p; p += b_off;
        if (m > 0) {
            PC_CUDA(ix, cudaMemcpyAsync((void *)d_q, q_xyz, (size_t)m * qs * sizeof(float), cudaMemcpyHostToDevice, st));
            PC_CUDA(ix, cudaMemcpyAsync((void *)d_r, range, (size_t)n_range * sizeof(double), cudaMemcpyHostToDevice, st));
        }
    }
    int64_t *d_cnt = (int64_t *)p; p += b_cnt;
    int64_t *d_tile = (int64_t *)p; p += b_tile;
    unsigned long long *d_long_count = (unsigned long long *)p;
    int64_t *d_long_list = (int64_t *)p + 1; p += b_long;
    pc_range_stage S;
    S.stage = (int32_t *)p; p += b_stage;
    S.stage_cap = (unsigned long long)stage_cap;
    S.pos = (int64_t *)p; p += b_pos;
    unsigned long long *d_todo_count = (unsigned long long *)p;
    int64_t *d_todo_list = (int64_t *)p + 1; p += b_todo;
    S.cursor = (unsigned long long *)p;
    const pc_range_stage no_stage = { nullptr, 0ull, nullptr, nullptr };

    int64_t total = 0;
    if (m == 0) {
        if (space == PC_HOST) out_offsets[0] = 0;
        else PC_CUDA(ix, cudaMemsetAsync(out_offsets, 0, sizeof(int64_t), st));
        return PC_OK;
    }
    pc_tree T = pc_tree_of(ix);
    const int cgrid = (int)((m + PC_RCOOP_WARPS - 1) / PC_RCOOP_WARPS);      // one warp per query
    if (capture) {
        PC_CUDA(ix, cudaMemsetAsync(S.cursor, 0, sizeof(unsigned long long), st));
        pc_range_coop_kernel<PC_RANGE_CAPTURE><<<cgrid, 32 * PC_RCOOP_WARPS, 0, st>>>(T, d_q, m, qs, d_r, range_is_scalar ? 1 : 0, d_cnt, nullptr, nullptr, nullptr, nullptr, S, nullptr, nullptr);
    } else {
        pc_range_coop_kernel<PC_RANGE_COUNT><<<cgrid, 32 * PC_RCOOP_WARPS, 0, st>>>(T, d_q, m, qs, d_r, range_is_scalar ? 1 : 0, d_cnt, nullptr, nullptr, nullptr, nullptr, no_stage, nullptr, nullptr);
    }
    pc_scan_tile_sums<<<(int)n_tiles, PC_SCAN_THREADS, 0, st>>>(d_cnt, m, d_tile);
    pc_scan_tile_offsets<<<1, PC_SCAN_THREADS, 0, st>>>(d_tile, n_tiles);
    pc_scan_write_offsets<<<(int)n_tiles, PC_SCAN_THREADS, 0, st>>>(d_cnt, m, d_tile, d_off);
    ix->launches += 4;
    PC_CHECK_LAUNCH(ix);
    // the total decides whether the lists fit: one 8-byte read-back
    PC_CUDA(ix, cudaMemcpyAsync(&total, d_off + m, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (space == PC_HOST)
        PC_CUDA(ix, cudaMemcpyAsync(out_offsets, d_off, (size_t)(m + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PC_CUDA(ix, cudaStreamSynchronize(st));
    tr.mark();            // 2: uploads + counting / capturing walk + scan + read-back
    if (!out_idx && cap == 0) return PC_OK;              // offsets only
    if (total > cap) return pc_fail(ix, PC_ECAP, "pc_range_batch: %lld hits exceed cap %lld", (long long)total, (long long)cap);
    if (total == 0) return PC_OK;

    int32_t *d_out = out_idx;
    if (space == PC_HOST) {
        // the lists are staged in the (idle) lane-0 int32 buffer
        pc_lane &L = ix->lane[0];
        if ((rc = pc_grow(ix, &L.d_i32, &L.i32_cap, total, total + total / 2)) != PC_OK) return rc;
        d_out = L.d_i32;
    }
    tr.mark();            // 3: list buffer
    PC_CUDA(ix, cudaMemsetAsync(d_long_count, 0, sizeof(unsigned long long), st));
    PC_CUDA(ix, cudaMemsetAsync(d_todo_count, 0, sizeof(unsigned long long), st));
    pc_range_place_kernel<<<(int)((m + PC_RPLACE_WARPS - 1) / PC_RPLACE_WARPS), 32 * PC_RPLACE_WARPS, 0, st>>>(m, d_off, S.stage, S.pos, d_out, d_todo_count, d_todo_list);
    ix->launches++;
    PC_CHECK_LAUNCH(ix);
    // the lists that were too long to park (each > PC_RCAP_HITS hits, so at most total / (PC_RCAP_HITS + 1) of them): second walk
    const int64_t max_todo = total / (PC_RCAP_HITS + 1) < m ? total / (PC_RCAP_HITS + 1) : m;
    if (max_todo > 0) {
        const int tgrid = (int)((max_todo + PC_RCOOP_WARPS - 1) / PC_RCOOP_WARPS);
        pc_range_coop_kernel<PC_RANGE_FILL><<<tgrid, 32 * PC_RCOOP_WARPS, 0, st>>>(T, d_q, m, qs, d_r, range_is_scalar ? 1 : 0, nullptr, d_off, d_out, d_long_count, d_long_list,
                                                                                  no_stage, d_todo_list, d_todo_count);
        ix->launches++;
        PC_CHECK_LAUNCH(ix);
    }
    if (max_todo > 0) {                  // only the lists of the fill pass can be longer than what a warp sorts
        static bool attr_set[64] = { false };
        if (ix->device >= 64 || !attr_set[ix->device]) {
            PC_CUDA(ix, cudaFuncSetAttribute(pc_range_sort_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_RLONG_CAP * (int)sizeof(uint32_t)));
            if (ix->device < 64) attr_set[ix->device] = true;
        }
        const int64_t max_long = max_todo;
        const int lgrid = (int)(max_long < (int64_t)ix->sm_count ? (max_long > 0 ? max_long : 1) : (int64_t)ix->sm_count);
        pc_range_sort_long_kernel<<<lgrid, PC_RLONG_THREADS, PC_RLONG_CAP * sizeof(uint32_t), st>>>(d_long_count, d_long_list, d_off, d_out);
        ix->launches++;
        PC_CHECK_LAUNCH(ix);
    }
    tr.mark();            // 4: place / fill / long-sort launches
    if (space == PC_HOST) {
        PC_CUDA(ix, cudaMemcpyAsync(out_idx, d_out, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        PC_CUDA(ix, cudaStreamSynchronize(st));
    }
    tr.mark();            // 5: lists home
    if (tr.limit > 0.0 && tr.n == 6 && tr.t[5] - tr.t[0] > tr.limit)
        fprintf(stderr, "pcindex: slow pc_range_batch %.1f ms (m %lld, hits %lld, long lists <= %lld): scratch %.1f, walk + scan %.1f, list buffer %.1f, launches %.1f, "
                "kernels + copy home %.1f\n", tr.t[5] - tr.t[0], (long long)m, (long long)total, (long long)max_todo,
                tr.t[1] - tr.t[0], tr.t[2] - tr.t[1], tr.t[3] - tr.t[2], tr.t[4] - tr.t[3], tr.t[5] - tr.t[4]);
    return PC_OK;
}

extern "C" int pc_sphere_gather(pc_index *ix, const double center[3], double radius, int space,
                                int32_t *out_idx, int64_t cap, int64_t *out_count)
{
    if (!ix) return PC_EINVAL;
    if (!center || !out_count || cap < 0 || (cap > 0 && !out_idx) || (space != PC_HOST && space != PC_DEVICE) || !(radius == radius))
        return pc_fail(ix, PC_EINVAL, "pc_sphere_gather: bad argument");
    PC_CUDA(ix, cudaSetDevice(ix->device));
    *out_count = 0;
    if (ix->n == 0) return PC_OK;
    cudaStream_t st = ix->stream;
    pc_lane &L = ix->lane[0];
    // only the real points: the tail of the last leaf repeats the last point
    const int64_t n = ix->n;
    const int64_t want = cap < n ? cap : n;
    int rc;
    if (want > L.sort_cap) {
        pc_pool_free(ix, L.keys_a); pc_pool_free(ix, L.keys_b); pc_pool_free(ix, L.vals_a); pc_pool_free(ix, L.vals_b);
        L.keys_a = L.keys_b = L.vals_a = L.vals_b = nullptr; L.sort_cap = 0;
        PC_CUDA(ix, pc_pool_alloc(ix, (void **)&L.keys_a, (size_t)want * 4));
        PC_CUDA(ix, pc_pool_alloc(ix, (void **)&L.keys_b, (size_t)want * 4));
        PC_CUDA(ix, pc_pool_alloc(ix, (void **)&L.vals_a, (size_t)want * 4));
        PC_CUDA(ix, pc_pool_alloc(ix, (void **)&L.vals_b, (size_t)want * 4));
        L.sort_cap = want;
    }
    if ((rc = pc_grow(ix, &L.tile_hist, &L.hist_cap, pc_sort_scratch_words(want > 0 ? want : 1, 16, 4))) != PC_OK) return rc;
    const float cx = (float)center[0], cy = (float)center[1], cz = (float)center[2];   // PCL searches with a float32 point
    const double r2 = radius * radius;
    float thr = (float)r2;
    thr = nextafterf(thr, INFINITY) * PC_THR_SLACK;
    if (!(thr < FLT_MAX)) thr = FLT_MAX;
    PC_CUDA(ix, cudaMemsetAsync(L.counter, 0, 2 * sizeof(unsigned long long), st));
    pc_sphere_gather_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(ix->points, n, cx, cy, cz, r2, thr, L.keys_a, (unsigned long long)want, L.counter + 1);
    ix->launches++;
    PC_CHECK_LAUNCH(ix);
    unsigned long long total = 0;
    PC_CUDA(ix, cudaMemcpyAsync(&total, L.counter + 1, sizeof total, cudaMemcpyDeviceToHost, st));
    PC_CUDA(ix, cudaStreamSynchronize(st));
    *out_count = (int64_t)total;
    if ((int64_t)total > cap) return cap == 0 && !out_idx ? PC_OK : pc_fail(ix, PC_ECAP, "pc_sphere_gather: %lld hits exceed cap %lld", (long long)total, (long long)cap);
    if (total == 0) return PC_OK;
    // ascending original index: sort the appended ids (values are not needed; the key buffer doubles as value buffer)
    int bits = 8;
    while (bits < 32 && ((unsigned long long)n >> bits) != 0) bits += 8;
    int which = pc_sort_pairs<uint32_t, 16>(ix, L.keys_a, L.vals_a, L.keys_b, L.vals_b, (int64_t)total, 0, bits, L.tile_hist, L.digit_total, st);
    PC_CHECK_LAUNCH(ix);
    const uint32_t *sorted = which ? L.keys_b : L.keys_a;
    PC_CUDA(ix, cudaMemcpyAsync(out_idx, sorted, (size_t)total * sizeof(int32_t), space == PC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
    if (space == PC_HOST) PC_CUDA(ix, cudaStreamSynchronize(st));
    return PC_OK;
}
