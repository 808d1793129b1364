// peer.cu — all-reduce of G1 partial sums over NVLink peer memory, fused into ONE kernel per prover phase.
//
// Point-range sharded MSM leaves one 96-byte Jacobian partial per GPU and commitment.  EC addition is
// not an NCCL reduction op, and the payload is tiny, so instead of an NCCL all-gather plus a separate summation
// kernel every rank runs k_g1_allreduce: it STORES its partials straight into every peer's mailbox through the
// NVLink-mapped (CUDA IPC) address, publishes a flag, waits for the flags of all peers in its own mailbox and adds
// the points with the 4-lane group law of quad.cuh.  One launch, no host round trip, no NCCL call on the data path.
#include "h2b_internal.cuh"
#include "quad.cuh"

namespace h2b {

static constexpr int PEER_MAX_RANKS = 16;
static constexpr int PEER_MAX_POINTS = 16;  // commitments per call (a prover phase has <= 4)

struct PeerMailbox {
    uint64_t data[2][PEER_MAX_RANKS][PEER_MAX_POINTS][12];  // [epoch parity][source rank][point][limb]
    uint64_t flags[2][PEER_MAX_RANKS];                      // epoch number published by each source rank
};

struct PeerPtrs {
    PeerMailbox* box[PEER_MAX_RANKS];
};

__device__ __forceinline__ void store_jacobian_xyzz(const XYZZ& p, uint64_t* out) {
    char* o = reinterpret_cast<char*>(out);
    if (p.is_identity()) {
        Fq::zero().store(o);
        Fq::one().store(o + 32);
        Fq::zero().store(o + 64);
        return;
    }
    Fq z = p.zz * p.zzz;
    Fq a = p.zz * p.zzz.sqr();
    (p.x * a).store(o);
    (p.y * a * p.zz.sqr()).store(o + 32);
    z.store(o + 64);
}

// Parity slots: epoch e uses slot e & 1.  A rank can only start epoch e+2 after it passed the wait of epoch e+1, and a
// peer publishes epoch e+1 only after its own kernel of epoch e has finished (stream order), so the data of epoch e
// is never overwritten while somebody still reads it.
__global__ void __launch_bounds__(256) k_g1_allreduce(PeerPtrs peers, int rank, int nranks, uint64_t epoch,
                                                      uint64_t* __restrict__ pts, int m) {
    const int par = (int)(epoch & 1);
    const int t = threadIdx.x;
    // 1. scatter my partials into every rank's mailbox (own one included), slot [parity][rank]
    const int words = m * 12;
    for (int idx = t; idx < words * nranks; idx += blockDim.x) {
        const int r = idx / words, w = idx % words;
        volatile uint64_t* dst = &peers.box[r]->data[par][rank][0][0];
        dst[w] = pts[w];
    }
    __threadfence_system();
    __syncthreads();
    if (t < nranks) {
        volatile uint64_t* f = &peers.box[t]->flags[par][rank];
        *f = epoch;
    }
    // 2. wait until every rank has published this epoch in MY mailbox
    if (t < nranks) {
        volatile uint64_t* f = &peers.box[rank]->flags[par][t];
        // bounded wait: a peer that never arrives (failed launch, dead process) must not hang the GPU; after ~10 s the
        // kernel traps, which surfaces as H2B_ERR_CUDA on the next call instead of a wedged device
        unsigned long long t0 = 0, now = 0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned spins = 0;
        while (*f != epoch) {
            if ((++spins & 0xfffu) == 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (now - t0 > 10000000000ull) __trap();
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    // 3. one lane-quad per commitment: sum over the source ranks
    const int q = t >> 2;
    if (q < m) {
        XYZZ acc = XYZZ::identity();
        for (int r = 0; r < nranks; r++) {
            const uint64_t* src = &peers.box[rank]->data[par][r][q][0];
            uint64_t tmp[12];
#pragma unroll
            for (int l = 0; l < 12; l++) tmp[l] = __ldcg(src + l);
            XYZZ p;  // Jacobian -> XYZZ
            Fq z;
#pragma unroll
            for (int l = 0; l < 4; l++) {
                p.x.l[2 * l] = (u32)tmp[l]; p.x.l[2 * l + 1] = (u32)(tmp[l] >> 32);
                p.y.l[2 * l] = (u32)tmp[4 + l]; p.y.l[2 * l + 1] = (u32)(tmp[4 + l] >> 32);
                z.l[2 * l] = (u32)tmp[8 + l]; z.l[2 * l + 1] = (u32)(tmp[8 + l] >> 32);
            }
            if (z.is_zero()) continue;
            p.zz = z.sqr();
            p.zzz = p.zz * z;
            quad_add_nl(acc, p);
        }
        if ((t & 3) == 0) store_jacobian_xyzz(acc, pts + 12 * q);
    }
}

struct PeerState {
    int rank = -1, nranks = 0;
    PeerMailbox* own = nullptr;
    PeerPtrs ptrs{};
    bool opened[PEER_MAX_RANKS] = {};
    bool connected = false;
    uint64_t epoch = 0;
};

void peer_create(h2b_ctx* ctx, int rank, int nranks, uint8_t* handle_out) {
    H2B_REQUIRE(nranks >= 1 && nranks <= PEER_MAX_RANKS && rank >= 0 && rank < nranks, "peer: bad rank / nranks");
    H2B_REQUIRE(!ctx->peer, "peer: mailbox already created");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    PeerState* st = new PeerState();
    st->rank = rank;
    st->nranks = nranks;
    ctx->peer = st;
    H2B_CUDA(cudaMalloc((void**)&st->own, sizeof(PeerMailbox)));
    H2B_CUDA(cudaMemset(st->own, 0, sizeof(PeerMailbox)));
    cudaIpcMemHandle_t h;
    H2B_CUDA(cudaIpcGetMemHandle(&h, st->own));
    memcpy(handle_out, &h, 64);
}

void peer_connect(h2b_ctx* ctx, const uint8_t* handles) {
    PeerState* st = (PeerState*)ctx->peer;
    H2B_REQUIRE(st, "peer: call h2b_peer_create first");
    for (int r = 0; r < st->nranks; r++) {
        if (r == st->rank) { st->ptrs.box[r] = st->own; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        H2B_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        st->ptrs.box[r] = (PeerMailbox*)p;
        st->opened[r] = true;
    }
    st->connected = true;
}

// One process, several devices: no IPC — every device enables peer access to the others and the mailboxes are addressed
// through their ordinary device pointers.
void peer_connect_local(const std::vector<h2b_ctx*>& members) {
    const int G = (int)members.size();
    H2B_REQUIRE(G >= 2 && G <= PEER_MAX_RANKS, "peer: 2..16 devices per group");
    for (int i = 0; i < G; i++) {
        H2B_REQUIRE(!members[i]->peer, "peer: mailbox already created");
        H2B_CUDA(cudaSetDevice(members[i]->device));
        PeerState* st = new PeerState();
        st->rank = i;
        st->nranks = G;
        members[i]->peer = st;
        H2B_CUDA(cudaMalloc((void**)&st->own, sizeof(PeerMailbox)));
        H2B_CUDA(cudaMemset(st->own, 0, sizeof(PeerMailbox)));
        for (int j = 0; j < G; j++) {
            if (j == i) continue;
            int can = 0;
            H2B_CUDA(cudaDeviceCanAccessPeer(&can, members[i]->device, members[j]->device));
            H2B_REQUIRE(can, "peer: the devices of the group cannot address each other (no NVLink / P2P path)");
            cudaError_t e = cudaDeviceEnablePeerAccess(members[j]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else H2B_CUDA(e);
        }
    }
    for (int i = 0; i < G; i++) {
        PeerState* st = (PeerState*)members[i]->peer;
        for (int j = 0; j < G; j++) st->ptrs.box[j] = ((PeerState*)members[j]->peer)->own;
        st->connected = true;
    }
    H2B_CUDA(cudaSetDevice(members[0]->device));
}

void peer_allreduce(h2b_ctx* ctx, void* d_points, size_t m) {
    PeerState* st = (PeerState*)ctx->peer;
    H2B_REQUIRE(st && st->connected, "peer: mailboxes are not connected");
    H2B_REQUIRE(m >= 1 && m <= (size_t)PEER_MAX_POINTS, "peer: 1..16 points per call");
    // the epoch only advances once the launch has been accepted: a failed launch must not desynchronise the ranks
    H2B_LAUNCH(ctx, k_g1_allreduce, 1, 256, 0, st->ptrs, st->rank, st->nranks, st->epoch + 1, (uint64_t*)d_points, (int)m);
    st->epoch++;
}
bool peer_connected(const h2b_ctx* ctx) {
    const PeerState* st = (const PeerState*)ctx->peer;
    return st && st->connected && st->nranks > 1;
}

void peer_destroy(h2b_ctx* ctx) {
    PeerState* st = (PeerState*)ctx->peer;
    if (!st) return;
    for (int r = 0; r < st->nranks; r++)
        if (st->opened[r]) cudaIpcCloseMemHandle(st->ptrs.box[r]);
    if (st->own) cudaFree(st->own);
    delete st;
    ctx->peer = nullptr;
}

}  // namespace h2b
