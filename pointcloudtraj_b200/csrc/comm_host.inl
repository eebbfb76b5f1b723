struct pc_comm { int rank, n_ranks, device; void *nccl; };
extern "C" int pc_comm_unique_id(char id[PC_NCCL_UNIQUE_ID_BYTES]) { (void)id; return PC_ENOTIMPL; }
extern "C" int pc_comm_init(pc_comm **out, int rank, int n_ranks, const char id[PC_NCCL_UNIQUE_ID_BYTES], int device)
{ (void)out; (void)rank; (void)n_ranks; (void)id; (void)device; return PC_ENOTIMPL; }
extern "C" void pc_comm_destroy(pc_comm *c) { (void)c; }
extern "C" int pc_index_broadcast(pc_index *ix, pc_comm *c, int root) { (void)c; (void)root; return pc_fail(ix, PC_ENOTIMPL, "pc_index_broadcast: not implemented yet"); }
