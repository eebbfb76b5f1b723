// comm_host.inl -- multi-GPU replication of a built index (included by pc_index.cu).
//
// One process per GPU.  The cloud index is read-only during queries, so the only exchange on the path is the
// replication of the index: rank `root` builds it, one ncclBroadcast over NVLink/NVSwitch copies the tree array (records +
// points, one contiguous span) to
// the other ranks' handles, and every rank then answers its own contiguous slice of the query batch
// (pc_shard_range) with no further collective.  NCCL is bound at run time with dlopen so that the library that a
// host framework (e.g. torch) already loaded is reused and libpcindex.so has no link-time NCCL dependency.

typedef struct { char internal[PC_NCCL_UNIQUE_ID_BYTES]; } pc_nccl_uid;   // layout of ncclUniqueId
typedef void *pc_nccl_comm_t;

struct pc_nccl_api {
    void *lib;
    int (*GetUniqueId)(pc_nccl_uid *);
    int (*CommInitRank)(pc_nccl_comm_t *, int, pc_nccl_uid, int);
    int (*CommDestroy)(pc_nccl_comm_t);
    int (*Broadcast)(const void *, void *, size_t, int /*ncclDataType_t*/, int, pc_nccl_comm_t, cudaStream_t);
    const char *(*GetErrorString)(int);
};

static pc_nccl_api g_nccl = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
static char g_comm_error[256] = "";

static int pc_nccl_load(void)
{
    if (g_nccl.lib) return PC_OK;
    const char *names[] = { getenv("PC_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { snprintf(g_comm_error, sizeof g_comm_error, "NCCL not found (dlopen libnccl.so.2): %s", dlerror()); return PC_ENCCL; }
    pc_nccl_api a;
    a.lib = h;
    a.GetUniqueId = (int (*)(pc_nccl_uid *))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(pc_nccl_comm_t *, int, pc_nccl_uid, int))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (int (*)(pc_nccl_comm_t))dlsym(h, "ncclCommDestroy");
    a.Broadcast = (int (*)(const void *, void *, size_t, int, int, pc_nccl_comm_t, cudaStream_t))dlsym(h, "ncclBroadcast");
    a.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Broadcast) {
        snprintf(g_comm_error, sizeof g_comm_error, "NCCL library lacks a required symbol");
        return PC_ENCCL;
    }
    g_nccl = a;
    return PC_OK;
}

struct pc_comm {
    int rank, n_ranks, device;
    pc_nccl_comm_t nccl;
};

#define PC_NCCL(ix, call)                                                                                     \
    do {                                                                                                      \
        int r_ = (call);                                                                                      \
        if (r_ != 0)                                                                                          \
            return pc_fail((ix), PC_ENCCL, "%s: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "NCCL error"); \
    } while (0)

extern "C" int pc_comm_unique_id(char id[PC_NCCL_UNIQUE_ID_BYTES])
{
    if (!id) return PC_EINVAL;
    int rc = pc_nccl_load();
    if (rc != PC_OK) { strncpy(g_create_error, g_comm_error, sizeof g_create_error - 1); return rc; }
    pc_nccl_uid u;
    PC_NCCL(nullptr, g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, PC_NCCL_UNIQUE_ID_BYTES);
    return PC_OK;
}

extern "C" int pc_comm_init(pc_comm **out, int rank, int n_ranks, const char id[PC_NCCL_UNIQUE_ID_BYTES], int device)
{
    if (!out || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return pc_fail(nullptr, PC_EINVAL, "pc_comm_init: bad argument");
    *out = nullptr;
    int rc = pc_nccl_load();
    if (rc != PC_OK) { strncpy(g_create_error, g_comm_error, sizeof g_create_error - 1); return rc; }
    PC_CUDA(nullptr, cudaSetDevice(device));
    pc_comm *c = new (std::nothrow) pc_comm();
    if (!c) return pc_fail(nullptr, PC_ENOMEM, "pc_comm_init: host allocation failed");
    c->rank = rank; c->n_ranks = n_ranks; c->device = device; c->nccl = nullptr;
    pc_nccl_uid u;
    memcpy(u.internal, id, PC_NCCL_UNIQUE_ID_BYTES);
    int r = g_nccl.CommInitRank(&c->nccl, n_ranks, u, rank);
    if (r != 0) {
        delete c;
        return pc_fail(nullptr, PC_ENCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    }
    *out = c;
    return PC_OK;
}

extern "C" void pc_comm_destroy(pc_comm *c)
{
    if (!c) return;
    if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
    delete c;
}

// header of a replicated index, broadcast ahead of the tree array so that receivers can size their arena
struct pc_bcast_header {
    int64_t n, n_nodes;
    uint32_t root, root_count;
    uint32_t bbox[6];
    uint32_t leaf;      // PC_LEAF of the sender (must match)
    uint32_t pad;
};

extern "C" int pc_index_broadcast(pc_index *ix, pc_comm *c, int root)
{
    if (!ix || !c || root < 0 || root >= c->n_ranks) return pc_fail(ix, PC_EINVAL, "pc_index_broadcast: bad argument");
    PC_CUDA(ix, cudaSetDevice(ix->device));
    cudaStream_t st = ix->stream;
    void *scr = nullptr;
    int rc = pc_scratch(ix, 256, &scr);
    if (rc != PC_OK) return rc;
    pc_bcast_header h;
    memset(&h, 0, sizeof h);
    if (c->rank == root) {
        h.n = ix->n; h.n_nodes = ix->n_nodes; h.root = ix->root; h.root_count = ix->root_count; h.leaf = PC_LEAF;
        if (ix->n > 0) PC_CUDA(ix, cudaMemcpyAsync(h.bbox, ix->d_bbox, sizeof h.bbox, cudaMemcpyDeviceToHost, st));
        PC_CUDA(ix, cudaStreamSynchronize(st));
        PC_CUDA(ix, cudaMemcpyAsync(scr, &h, sizeof h, cudaMemcpyHostToDevice, st));
    }
    PC_NCCL(ix, g_nccl.Broadcast(scr, scr, sizeof h, 0 /* ncclInt8 */, root, c->nccl, st));
    PC_CUDA(ix, cudaMemcpyAsync(&h, scr, sizeof h, cudaMemcpyDeviceToHost, st));
    PC_CUDA(ix, cudaStreamSynchronize(st));
    if (h.leaf != PC_LEAF) return pc_fail(ix, PC_EINVAL, "pc_index_broadcast: ranks were built with different PC_LEAF");
    if (c->rank != root) {
        if (h.n > ix->cap) {
            if ((rc = pc_reserve_cloud(ix, h.n)) != PC_OK) return rc;
        }
        ix->n = h.n; ix->n_nodes = h.n_nodes; ix->root = h.root; ix->root_count = h.root_count; ix->build_timed = false;
        ix->points = ix->tree + 4 * h.n_nodes;
        if (h.n > 0) PC_CUDA(ix, cudaMemcpyAsync(ix->d_bbox, h.bbox, sizeof h.bbox, cudaMemcpyHostToDevice, st));
        memcpy(ix->h_bbox, h.bbox, sizeof h.bbox);
        ix->bbox_from_bcast = h.n > 0;
    }
    if (h.n > 0) {
        // records [0, 4 n_nodes), the points (+ PC_LEAF pad copies) and the seeds behind them are one contiguous span of the tree array
        const size_t bytes = (size_t)(4 * h.n_nodes + h.n + PC_LEAF) * sizeof(float4) + PC_SEED_WORDS * sizeof(uint32_t);   // + the seeds
        if ((int64_t)(bytes / sizeof(float4)) > ix->tree_cap) return pc_fail(ix, PC_ENOMEM, "pc_index_broadcast: arena too small");
        PC_NCCL(ix, g_nccl.Broadcast(ix->tree, ix->tree, bytes, 0 /* ncclInt8 */, root, c->nccl, st));
    }
    PC_CUDA(ix, cudaEventRecord(ix->ev_ready, st));
    return PC_OK;
}
