// Halo exchange and small all-reduces of the element-partitioned path, one C call each (sm_100a + NCCL over NVLink).
//
// A halo plan holds, per rank, the concatenated send / receive index lists of all neighbours. One call packs the owned
// boundary values (k_pack), posts every ncclSend / ncclRecv of the step inside one NCCL group on the caller's stream and
// unpacks (copy: refresh ghosts; add: sum the neighbours' partial contributions) — ~20 us of host time instead of the
// ~90 us a batched torch isend/irecv costs, which is what bounded the distributed V-cycle in the first version.
// NCCL is resolved at run time from the libnccl.so.2 torch has already loaded (no link-time dependency).
//
// The reference has no counterpart (no distributed code); this implements SURVEY 8(e).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../include/opencmp_b200.h"
#include "ocmp_common.cuh"

namespace {
typedef struct { char internal[128]; } ncclUniqueId;
typedef void* ncclComm_t;
typedef int ncclResult_t;
const int NCCL_FLOAT64 = 8, NCCL_SUM = 0;
struct Nccl {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok = false;
} N;
ncclComm_t g_comm = nullptr;
int g_nranks = 1, g_rank = 0;

bool load_nccl() {
    if (N.ok) return true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (!h) return false;
#define SYM(field, name) *(void**)(&N.field) = dlsym(h, name); if (!N.field) return false;
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    N.ok = true;
    return true;
}

struct HaloPlan {
    int nnbr;
    std::vector<int> peers, send_cnt, recv_cnt;
    const int* send_idx;   // device, concatenated in peer order
    const int* recv_idx;
    int nsend, nrecv;
    double* sendbuf;       // device, nsend
    double* recvbuf;       // device, nrecv
};
std::vector<HaloPlan> g_plans;

__global__ void __launch_bounds__(256) k_pack(int n, const int* __restrict__ idx, const double* __restrict__ x,
                                              double* __restrict__ buf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = x[idx[i]];
}
__global__ void __launch_bounds__(256) k_unpack(int n, const int* __restrict__ idx, const double* __restrict__ buf,
                                                double* __restrict__ x, int add) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (add) atomicAdd(x + idx[i], buf[i]);      // a DOF shared with two neighbours receives two contributions
    else x[idx[i]] = buf[i];
}
}  // namespace

extern "C" int ocmp_comm_unique_id(void* out128) {
    if (!load_nccl()) return ocmp_fail(-20, "libnccl.so.2 not found");
    ncclUniqueId id;
    const ncclResult_t r = N.GetUniqueId(&id);
    if (r) return ocmp_fail(-21, N.GetErrorString(r));
    memcpy(out128, &id, 128);
    return 0;
}

extern "C" int ocmp_comm_init(const void* id128, int nranks, int rank) {
    if (!load_nccl()) return ocmp_fail(-20, "libnccl.so.2 not found");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    const ncclResult_t r = N.CommInitRank(&g_comm, nranks, id, rank);
    if (r) return ocmp_fail(-21, N.GetErrorString(r));
    g_nranks = nranks; g_rank = rank;
    return 0;
}

extern "C" int ocmp_halo_plan(int nnbr, const int* peers, const int* send_cnt, const int* recv_cnt,
                              const int* send_idx_dev, const int* recv_idx_dev, double* sendbuf_dev,
                              double* recvbuf_dev) {
    HaloPlan p;
    p.nnbr = nnbr;
    p.peers.assign(peers, peers + nnbr);
    p.send_cnt.assign(send_cnt, send_cnt + nnbr);
    p.recv_cnt.assign(recv_cnt, recv_cnt + nnbr);
    p.send_idx = send_idx_dev; p.recv_idx = recv_idx_dev;
    p.nsend = p.nrecv = 0;
    for (int i = 0; i < nnbr; ++i) { p.nsend += send_cnt[i]; p.nrecv += recv_cnt[i]; }
    p.sendbuf = sendbuf_dev; p.recvbuf = recvbuf_dev;
    g_plans.push_back(p);
    return (int)g_plans.size() - 1;
}

extern "C" int ocmp_halo_run(int plan, double* x, int add, void* stream) {
    if (plan < 0 || plan >= (int)g_plans.size()) return ocmp_fail(-1, "bad halo plan");
    if (!g_comm) return ocmp_fail(-22, "ocmp_comm_init was not called");
    const HaloPlan& p = g_plans[plan];
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(PROF_HALO, st);
    if (p.nsend > 0) k_pack<<<(p.nsend + 255) / 256, 256, 0, st>>>(p.nsend, p.send_idx, x, p.sendbuf);
    N.GroupStart();
    int so = 0, ro = 0;
    for (int i = 0; i < p.nnbr; ++i) {
        if (p.send_cnt[i]) N.Send(p.sendbuf + so, p.send_cnt[i], NCCL_FLOAT64, p.peers[i], g_comm, st);
        if (p.recv_cnt[i]) N.Recv(p.recvbuf + ro, p.recv_cnt[i], NCCL_FLOAT64, p.peers[i], g_comm, st);
        so += p.send_cnt[i]; ro += p.recv_cnt[i];
    }
    const ncclResult_t r = N.GroupEnd();
    if (r) return ocmp_fail(-21, N.GetErrorString(r));
    if (p.nrecv > 0) k_unpack<<<(p.nrecv + 255) / 256, 256, 0, st>>>(p.nrecv, p.recv_idx, p.recvbuf, x, add);
    return ocmp_check("ocmp_halo_run");
}

extern "C" int ocmp_allreduce_sum(double* buf_dev, int n, void* stream) {
    if (!g_comm) return ocmp_fail(-22, "ocmp_comm_init was not called");
    const ncclResult_t r = N.AllReduce(buf_dev, buf_dev, n, NCCL_FLOAT64, NCCL_SUM, g_comm, (cudaStream_t)stream);
    if (r) return ocmp_fail(-21, N.GetErrorString(r));
    return 0;
}
