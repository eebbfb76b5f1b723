// sampler_host.inl -- host side of pc_sample_batch / pc_expand_batch (included by pc_index.cu).

static int pc_make_sampler_dev(pc_index *ix, const pc_sampler *s, pc_sampler_dev *S)
{
    if (!s) return pc_fail(ix, PC_EINVAL, "pc_sampler: null");
    if (s->engine_state == 0u || s->engine_state >= PC_LCG_M)
        return pc_fail(ix, PC_EINVAL, "pc_sampler: engine_state %u is not a minstd_rand0 state (1 .. 2^31-2)", s->engine_state);
    if (!(s->goal_ratio == s->goal_ratio) || !(s->inlier_ratio == s->inlier_ratio)) return pc_fail(ix, PC_EINVAL, "pc_sampler: NaN ratio");
    S->state = s->engine_state;
    S->goal_ratio = s->goal_ratio;
    S->inlier_sum = s->goal_ratio + s->inlier_ratio;             // the reference compares with (goal_ratio + inlier_ratio)
    for (int a = 0; a < 3; a++) {
        S->end_pt[a] = s->end_pt[a];
        S->lo[a] = s->lo[a]; S->span[a] = s->hi[a] - s->lo[a];            // uniform_real_distribution: b - a, then * and +
        S->in_lo[a] = s->in_lo[a]; S->in_span[a] = s->in_hi[a] - s->in_lo[a];
    }
    return PC_OK;
}

#define PC_EXPAND_CHUNK ((int64_t)1 << 21)      // samples per pipelined chunk of a large expansion batch
#define PC_EXPAND_MAX_CHUNKS 512

// device buffers of one sample stream of k samples
struct pc_sample_bufs { uint32_t *masks; pc_jump *tile_map; uint2 *tile_entry; uint32_t *state; int64_t n_tiles; };

static int64_t pc_sample_tiles(int64_t k) { return (4 * k + 1 + PC_SMP_TILE - 1) / PC_SMP_TILE; }   // sample k starts at a position <= 4k
static int64_t pc_sample_scratch_bytes(int64_t k)
{
    const int64_t t = pc_sample_tiles(k);
    return pc_align_up(t * PC_SMP_THREADS * (int64_t)sizeof(uint32_t), 256) + pc_align_up(t * (int64_t)sizeof(pc_jump), 256) +
           pc_align_up(t * (int64_t)sizeof(uint2), 256) + 256;
}
static char *pc_sample_carve(char *p, int64_t k, pc_sample_bufs *B)
{
    const int64_t t = pc_sample_tiles(k);
    B->n_tiles = t;
    B->masks = (uint32_t *)p; p += pc_align_up(t * PC_SMP_THREADS * (int64_t)sizeof(uint32_t), 256);
    B->tile_map = (pc_jump *)p; p += pc_align_up(t * (int64_t)sizeof(pc_jump), 256);
    B->tile_entry = (uint2 *)p; p += pc_align_up(t * (int64_t)sizeof(uint2), 256);
    B->state = (uint32_t *)p; p += 256;
    return p;
}

// the three kernels of the stream; d_state receives the engine state behind the k-th sample
static int pc_sample_launch(pc_index *ix, const pc_sampler_dev &S, int64_t k, const pc_sample_bufs &B, double *d_xyz, float4 *d_q, cudaStream_t st)
{
    if (4 * k + 1 <= (int64_t)PC_SMP_SCAN_THREADS * PC_SMP_CHUNK) {        // a planner-sized batch: one CTA does it all
        pc_sample_emit_kernel<true><<<1, PC_SMP_SCAN_THREADS, 0, st>>>(S, nullptr, nullptr, (uint64_t)k, d_xyz, d_q, B.state);
        ix->launches++;
        PC_CHECK_LAUNCH(ix);
        return PC_OK;
    }
    pc_sample_mask_kernel<<<(int)B.n_tiles, PC_SMP_THREADS, 0, st>>>(S.state, S.goal_ratio, B.masks, B.tile_map);
    pc_sample_tile_kernel<<<1, PC_SMP_SCAN_THREADS, 0, st>>>(B.tile_map, B.n_tiles, B.tile_entry);
    pc_sample_emit_kernel<false><<<(int)B.n_tiles, PC_SMP_THREADS, 0, st>>>(S, B.masks, B.tile_entry, (uint64_t)k, d_xyz, d_q, B.state);
    ix->launches += 3;
    PC_CHECK_LAUNCH(ix);
    return PC_OK;
}

extern "C" int pc_sample_batch(pc_index *ix, const pc_sampler *sampler, int64_t k, int space, double *out_xyz, uint32_t *out_engine_state)
{
    if (!ix) return PC_EINVAL;
    if (k < 0 || k > ((int64_t)1 << 29) || (k > 0 && !out_xyz) || (space != PC_HOST && space != PC_DEVICE))
        return pc_fail(ix, PC_EINVAL, "pc_sample_batch: bad argument (k=%lld space=%d)", (long long)k, space);
    pc_sampler_dev S;
    int rc = pc_make_sampler_dev(ix, sampler, &S);
    if (rc != PC_OK) return rc;
    if (out_engine_state) *out_engine_state = sampler->engine_state;
    if (k == 0) return PC_OK;
    PC_CUDA(ix, cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const int64_t b_xyz = space == PC_HOST ? pc_align_up(3 * k * (int64_t)sizeof(double), 256) : 0;
    void *base = nullptr;
    if ((rc = pc_scratch(ix, pc_sample_scratch_bytes(k) + b_xyz, &base)) != PC_OK) return rc;
    pc_sample_bufs B;
    char *p = pc_sample_carve((char *)base, k, &B);
    double *d_xyz = space == PC_HOST ? (double *)p : out_xyz;
    if ((rc = pc_sample_launch(ix, S, k, B, d_xyz, nullptr, st)) != PC_OK) return rc;
    if (space == PC_HOST) PC_CUDA(ix, cudaMemcpyAsync(out_xyz, d_xyz, (size_t)(3 * k) * sizeof(double), cudaMemcpyDeviceToHost, st));
    // the engine state is a host value in both spaces (the caller's engine continues from it): one 4-byte read-back
    uint32_t state = 0;
    PC_CUDA(ix, cudaMemcpyAsync(&state, B.state, sizeof state, cudaMemcpyDeviceToHost, st));
    PC_CUDA(ix, cudaStreamSynchronize(st));
    if (out_engine_state) *out_engine_state = state;
    return PC_OK;
}

extern "C" int pc_expand_batch(pc_index *cloud, pc_index *nodes, const pc_node_set *set, const pc_sampler *sampler,
                               const pc_radius_params *params, double z_l, double safety_margin, int64_t k,
                               pc_candidate *out, int64_t cap, int64_t *out_count, uint32_t *out_engine_state)
{
    pc_index *ix = cloud;
    if (!ix) return PC_EINVAL;
    if (!nodes || nodes == cloud || !set || set->n < 1 || !set->coord || !set->radius || !set->valid || k < 0 || k > ((int64_t)1 << 29) ||
        cap < 0 || (cap > 0 && !out) || !out_count)
        return pc_fail(ix, PC_EINVAL, "pc_expand_batch: bad argument");
    if (nodes->device != ix->device) return pc_fail(ix, PC_EINVAL, "pc_expand_batch: the two handles live on different devices");
    if (set->n > ((int64_t)1 << 31) - 16) return pc_fail(ix, PC_EINVAL, "pc_expand_batch: too many nodes");
    pc_sampler_dev S;
    int rc = pc_make_sampler_dev(ix, sampler, &S);
    if (rc != PC_OK) return rc;
    pc_qargs RA;
    memset(&RA, 0, sizeof RA);
    RA.kind = PC_Q_RADIUS; RA.flags = PC_QUERY_AUTO;
    if ((rc = pc_make_radius_dev(ix, params, PC_RADIUS_BOUNDED, &RA.R)) != PC_OK) return rc;
    *out_count = 0;
    if (out_engine_state) *out_engine_state = sampler->engine_state;
    if (k == 0) return PC_OK;
    PC_CUDA(ix, cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    const int64_t n = set->n;

    // one slice of the cloud handle's scratch: node set | sample stream scratch | per-sample arrays | candidates
    const int64_t b_coord = pc_align_up(3 * n * (int64_t)sizeof(double), 256), b_rad = pc_align_up(n * (int64_t)sizeof(float), 256);
    const int64_t b_valid = pc_align_up(n, 256), b_pos = pc_align_up(n * (int64_t)sizeof(float4), 256);
    const int64_t b_xyz = pc_align_up(3 * k * (int64_t)sizeof(double), 256), b_q = pc_align_up(k * (int64_t)sizeof(float4), 256);
    const int64_t b_nn = pc_align_up(k * (int64_t)sizeof(int32_t), 256), b_r = pc_align_up(k * (int64_t)sizeof(float), 256), b_ok = pc_align_up(k, 256);
    if ((k + PC_EXPAND_CHUNK - 1) / PC_EXPAND_CHUNK > PC_EXPAND_MAX_CHUNKS) return pc_fail(ix, PC_EINVAL, "pc_expand_batch: batch too large");
    const int64_t n_ctile = (PC_EXPAND_CHUNK + PC_EXPAND_CHUNK / 2 + PC_CAND_TILE - 1) / PC_CAND_TILE;
    const int64_t b_ctile = pc_align_up(n_ctile * (int64_t)sizeof(uint32_t), 256);
    const int64_t b_total = pc_align_up((PC_EXPAND_MAX_CHUNKS + 1) * (int64_t)sizeof(unsigned long long), 256);
    const int64_t want_out = cap < k ? cap : k;
    const int64_t b_out = pc_align_up(want_out * (int64_t)sizeof(pc_candidate_dev), 256);
    void *base = nullptr;
    if ((rc = pc_scratch(ix, b_coord + b_rad + b_valid + b_pos + pc_sample_scratch_bytes(k) + b_xyz + b_q + b_nn + b_r + b_ok + b_ctile + b_total + b_out, &base)) != PC_OK) return rc;
    char *p = (char *)base;
    double *d_coord = (double *)p; p += b_coord;
    float *d_rad = (float *)p; p += b_rad;
    uint8_t *d_valid = (uint8_t *)p; p += b_valid;
    float4 *d_pos = (float4 *)p; p += b_pos;
    pc_sample_bufs B;
    p = pc_sample_carve(p, k, &B);
    double *d_xyz = (double *)p; p += b_xyz;
    float4 *d_q = (float4 *)p; p += b_q;
    int32_t *d_nn = (int32_t *)p; p += b_nn;
    float *d_r = (float *)p; p += b_r;
    uint8_t *d_ok = (uint8_t *)p; p += b_ok;
    uint32_t *d_ctile = (uint32_t *)p; p += b_ctile;
    unsigned long long *d_total = (unsigned long long *)p; p += b_total;      // running candidate count behind every chunk
    pc_candidate_dev *d_out = (pc_candidate_dev *)p;

    // the frozen node set: a few bytes per node (the planner's tree has 10^3 .. 10^5 nodes), then its index
    if (b_coord + b_rad + b_valid <= PC_CLR_STAGE_BYTES) {
        // a planner-sized node set: one copy from the pinned staging block instead of three pageable ones (~10 us each)
        if (!ix->h_stage) PC_CUDA(ix, cudaHostAlloc((void **)&ix->h_stage, 2 * PC_CLR_STAGE_BYTES, cudaHostAllocDefault));
        memcpy(ix->h_stage, set->coord, (size_t)(3 * n) * sizeof(double));
        memcpy(ix->h_stage + b_coord, set->radius, (size_t)n * sizeof(float));
        memcpy(ix->h_stage + b_coord + b_rad, set->valid, (size_t)n);
        PC_CUDA(ix, cudaMemcpyAsync(d_coord, ix->h_stage, (size_t)(b_coord + b_rad + b_valid), cudaMemcpyHostToDevice, st));
    } else {
        PC_CUDA(ix, cudaMemcpyAsync(d_coord, set->coord, (size_t)(3 * n) * sizeof(double), cudaMemcpyHostToDevice, st));
        PC_CUDA(ix, cudaMemcpyAsync(d_rad, set->radius, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, st));
        PC_CUDA(ix, cudaMemcpyAsync(d_valid, set->valid, (size_t)n, cudaMemcpyHostToDevice, st));
    }
    pc_node_pos_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(d_coord, n, d_pos);
    ix->launches++;
    if ((rc = pc_sample_launch(ix, S, k, B, d_xyz, d_q, st)) != PC_OK) return rc;
    PC_CUDA(ix, cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st));
    // a planner-sized batch against a planner-sized node set needs no index: exact brute force, one launch (sampler_kernels.cuh)
    const bool brute = k * n <= PC_BRUTE_MAX_PAIRS;
    if (!brute) {
        // the node index and the nearest-vertex batches run on the nodes handle (its own stream), ordered by events
        PC_CUDA(ix, cudaEventRecord(ix->ev_in, st));
        PC_CUDA(ix, cudaStreamWaitEvent(nodes->stream, ix->ev_in, 0));
        if ((rc = pc_index_build(nodes, (const float *)d_pos, n, 4, PC_DEVICE)) != PC_OK) return pc_fail(ix, rc, "pc_expand_batch: node index: %s", nodes->err);
    }

    // Large batches go through in chunks: while chunk c is searched, the candidates of chunk c - 1 travel to the host on a
    // side stream, and the nearest-vertex search of chunk c + 1 (nodes handle's stream) overlaps the radius search of chunk c.
    const int64_t chunk = k > PC_EXPAND_CHUNK + PC_EXPAND_CHUNK / 2 ? PC_EXPAND_CHUNK : k;
    const int64_t n_chunks = (k + chunk - 1) / chunk;
    if (!ix->h_totals) {
        PC_CUDA(ix, cudaHostAlloc((void **)&ix->h_totals, (PC_EXPAND_MAX_CHUNKS + 2) * sizeof(unsigned long long), cudaHostAllocDefault));
        for (int i = 0; i < 2; i++) PC_CUDA(ix, cudaEventCreateWithFlags(&ix->ev_chunk[i], cudaEventDisableTiming));
    }
    cudaStream_t copy_st = ix->lane[1].stream;
    uint32_t state = 0;
    unsigned long long copied = 0, total = 0;
    const bool eager = n_chunks == 1 && want_out * (int64_t)sizeof(pc_candidate_dev) <= ((int64_t)1 << 20);
    auto drain = [&](int64_t c) -> int {          // chunk c is done on the device: send its candidates home
        PC_CUDA(ix, cudaEventSynchronize(ix->ev_chunk[c & 1]));
        total = ix->h_totals[c + 1];
        const unsigned long long upto = total < (unsigned long long)cap ? total : (unsigned long long)cap;
        if (upto > copied && !eager)
            PC_CUDA(ix, cudaMemcpyAsync(out + copied, d_out + copied, (size_t)(upto - copied) * sizeof(pc_candidate_dev), cudaMemcpyDeviceToHost, copy_st));
        if (upto > copied) copied = upto;
        return PC_OK;
    };
    for (int64_t c = 0; c < n_chunks; c++) {
        const int64_t off = c * chunk, kc = k - off < chunk ? k - off : chunk;
        if (brute) {
            pc_nearest_brute_kernel<<<(int)((kc * 32 + 255) / 256), 256, 0, st>>>(d_pos, n, d_q + off, kc, d_nn + off);
            ix->launches++;
        } else {
            pc_qargs NA;
            memset(&NA, 0, sizeof NA);
            NA.kind = PC_Q_NEAREST; NA.flags = PC_QUERY_AUTO;
            if ((rc = pc_run_batch(nodes, nodes->lane[0], NA, (const float *)(d_q + off), kc, 4, d_nn + off, nullptr)) != PC_OK)
                return pc_fail(ix, rc, "pc_expand_batch: nearest vertex: %s", nodes->err);
            PC_CUDA(ix, cudaEventRecord(nodes->ev_in, nodes->stream));
            PC_CUDA(ix, cudaStreamWaitEvent(st, nodes->ev_in, 0));
        }
        pc_steer_kernel<<<(int)((kc + 255) / 256), 256, 0, st>>>(d_xyz + 3 * off, d_q + off, d_nn + off, kc, d_coord, d_rad, d_valid, d_ok + off);
        ix->launches++;
        PC_CHECK_LAUNCH(ix);
        // the steering overwrote this chunk's samples with the centres: the nodes stream must not run ahead into them -- it does
        // not, its next chunk reads other entries of d_q.  radiusSearch for every centre (a sample without a valid nearest
        // vertex keeps its own position: answered and dropped)
        if ((rc = pc_run_batch(ix, ix->lane[0], RA, (const float *)(d_q + off), kc, 4, nullptr, d_r + off)) != PC_OK) return rc;
        const int64_t n_ct = (kc + PC_CAND_TILE - 1) / PC_CAND_TILE;
        if (kc <= PC_CAND_SMALL_MAX) {
            pc_cand_small_kernel<<<1, 1024, 0, st>>>(d_xyz + 3 * off, d_r + off, d_ok + off, d_nn + off, kc, z_l, safety_margin, d_total + c, d_total + c + 1,
                                                     d_out, (uint64_t)want_out);
            ix->launches++;
        } else {
            pc_cand_count_kernel<<<(int)n_ct, PC_CAND_THREADS, 0, st>>>(d_xyz + 3 * off, d_r + off, d_ok + off, kc, z_l, safety_margin, d_ctile);
            pc_cand_scan_kernel<<<1, 1024, 0, st>>>(d_ctile, n_ct, d_total + c, d_total + c + 1);
            pc_cand_write_kernel<<<(int)n_ct, PC_CAND_THREADS, 0, st>>>(d_xyz + 3 * off, d_r + off, d_ok + off, d_nn + off, kc, z_l, safety_margin, d_ctile, d_out, (uint64_t)want_out);
            ix->launches += 3;
        }
        PC_CHECK_LAUNCH(ix);
        PC_CUDA(ix, cudaMemcpyAsync(ix->h_totals + c + 1, d_total + c + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        if (c == n_chunks - 1) {
            PC_CUDA(ix, cudaMemcpyAsync(ix->h_totals + PC_EXPAND_MAX_CHUNKS + 1, B.state, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            // a planner-sized batch comes back with its count in one round trip
            if (eager && want_out > 0) PC_CUDA(ix, cudaMemcpyAsync(out, d_out, (size_t)want_out * sizeof(pc_candidate_dev), cudaMemcpyDeviceToHost, st));
        }
        PC_CUDA(ix, cudaEventRecord(ix->ev_chunk[c & 1], st));
        if (c > 0 && (rc = drain(c - 1)) != PC_OK) return rc;
    }
    if ((rc = drain(n_chunks - 1)) != PC_OK) return rc;
    PC_CUDA(ix, cudaStreamSynchronize(st));
    if (!eager) PC_CUDA(ix, cudaStreamSynchronize(copy_st));
    state = *(const uint32_t *)(ix->h_totals + PC_EXPAND_MAX_CHUNKS + 1);
    *out_count = (int64_t)total;
    if (out_engine_state) *out_engine_state = state;
    if ((int64_t)total > cap) return pc_fail(ix, PC_ECAP, "pc_expand_batch: %lld candidates exceed cap %lld", (long long)total, (long long)cap);
    return PC_OK;
}
