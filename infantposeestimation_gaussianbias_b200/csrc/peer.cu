// peer.cu — mailboxes for the scalar exchanges of a batch-sharded job (one process per GPU).
//
// The sharded loss needs two exchanges per step: the batch normalisers
// [sum w, sum w_i w_j] before the tile kernel and the seven loss scalars after it.  Instead of two
// NCCL all-reduces (~15-30 us each of launch + protocol latency on a 0.3 ms step) the kernels that
// produce those scalars write them straight into every peer's mailbox over NVLink / NVSwitch
// (cudaIpc-mapped device memory) and read the peers' values from their own mailbox:
// denoms_kernel and finalize_kernel in loss.cu.  This file is the set-up: allocate the local
// mailbox, export its IPC handle, map the peers'.  The handles travel through whatever the host
// program already has (torch.distributed all_gather in sharded.py).
#include "loss_common.cuh"

namespace gbc {

int peer_create(int rank, int world, void** ctx_out, unsigned char* handle_out) {
    if (!ctx_out || !handle_out) return fail(GBCODEC_ERR_NULL_POINTER, "peer_create: NULL pointer");
    if (world < 1 || world > GBCODEC_MAX_PEERS || rank < 0 || rank >= world)
        return fail(GBCODEC_ERR_BAD_ARGUMENT, "peer_create: rank %d of %d (at most %d ranks)", rank, world, GBCODEC_MAX_PEERS);
    static_assert(sizeof(cudaIpcMemHandle_t) <= GBCODEC_PEER_HANDLE_BYTES, "IPC handle does not fit");
    PeerCtx* pc = new PeerCtx();
    memset(pc, 0, sizeof(*pc));
    pc->view.rank = rank; pc->view.world = world;
    cudaError_t e = cudaGetDevice(&pc->device);
    PeerMail* mail = nullptr;
    pc->view.timeout_ns = 120ull * 1000000000ull;
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&pc->h_failed), sizeof(unsigned int), cudaHostAllocMapped);
    if (e == cudaSuccess) { *pc->h_failed = 0u; e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&pc->view.failed), pc->h_failed, 0); }
    if (e == cudaSuccess) e = cudaMalloc(&mail, sizeof(PeerMail));
    if (e == cudaSuccess) e = cudaMemset(mail, 0, sizeof(PeerMail));
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&pc->handle, mail);
    if (e != cudaSuccess) {
        if (mail) cudaFree(mail);
        if (pc->h_failed) cudaFreeHost(pc->h_failed);
        delete pc;
        return fail(GBCODEC_ERR_CUDA, "peer_create: %s", cudaGetErrorString(e));
    }
    pc->view.mail[rank] = mail;
    memset(handle_out, 0, GBCODEC_PEER_HANDLE_BYTES);
    memcpy(handle_out, &pc->handle, sizeof(pc->handle));
    pc->connected = world == 1;
    *ctx_out = pc;
    return GBCODEC_OK;
}

int peer_connect(void* ctx, const unsigned char* handles) {
    if (!ctx || !handles) return fail(GBCODEC_ERR_NULL_POINTER, "peer_connect: NULL pointer");
    PeerCtx* pc = reinterpret_cast<PeerCtx*>(ctx);
    for (int r = 0; r < pc->view.world; ++r) {
        if (r == pc->view.rank || pc->view.mail[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * GBCODEC_PEER_HANDLE_BYTES, sizeof(h));
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "peer_connect: rank %d cannot map the mailbox of rank %d: %s", pc->view.rank, r, cudaGetErrorString(e));
        pc->view.mail[r] = reinterpret_cast<PeerMail*>(p);
    }
    pc->connected = 1;
    return GBCODEC_OK;
}

int peer_status(void* ctx, int* timeouts) {
    if (!ctx || !timeouts) return fail(GBCODEC_ERR_NULL_POINTER, "peer_status: NULL pointer");
    PeerCtx* pc = reinterpret_cast<PeerCtx*>(ctx);
    unsigned int t = 0;
    const cudaError_t e = cudaMemcpy(&t, &pc->view.mail[pc->view.rank]->timeouts, sizeof(t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(GBCODEC_ERR_CUDA, "peer_status: %s", cudaGetErrorString(e));
    *timeouts = (int)t;
    return GBCODEC_OK;
}

int peer_set_timeout(void* ctx, double seconds) {
    if (!ctx) return fail(GBCODEC_ERR_NULL_POINTER, "peer_set_timeout: NULL context");
    if (!(seconds > 0.0) || seconds > 86400.0) return fail(GBCODEC_ERR_BAD_ARGUMENT, "peer_set_timeout: %g s", seconds);
    reinterpret_cast<PeerCtx*>(ctx)->view.timeout_ns = (unsigned long long)(seconds * 1e9);
    return GBCODEC_OK;
}

int peer_destroy(void* ctx) {
    if (!ctx) return GBCODEC_OK;
    PeerCtx* pc = reinterpret_cast<PeerCtx*>(ctx);
    if (pc->h_failed) cudaFreeHost(pc->h_failed);
    for (int r = 0; r < pc->view.world; ++r) {
        if (!pc->view.mail[r]) continue;
        if (r == pc->view.rank) cudaFree(pc->view.mail[r]);
        else cudaIpcCloseMemHandle(pc->view.mail[r]);
    }
    delete pc;
    return GBCODEC_OK;
}

}  // namespace gbc
