// Internal interface of the slab communicator (ny_comm.cu).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include "../../include/nyles_b200.h"

struct ny_ctx;

// Face exchange over NVLink peer memory (ny_comm.cu, "P2P halo exchange"): every rank owns one buffer of
// receive slots that its two slab neighbours map through CUDA IPC and write into directly.
constexpr int NY_P2P_SLOTS = 4;
struct ny_p2p {
    int state;                          // 0: not set up yet, 1: in use, -1: unavailable (NCCL send/recv is used)
    size_t slot_bytes;                  // capacity of one receive slot
    unsigned char* local;               // [2 directions][NY_P2P_SLOTS][slot_bytes], then the arrival flags
    unsigned long long* flags;          // local arrival flags [2][NY_P2P_SLOTS] (written by the neighbours)
    unsigned int* counters;             // [NY_P2P_SLOTS] CTAs of a push that have finished their share
    unsigned char* peer[2];             // the buffer of the rank below / above as mapped into this process
    int peer_rank[2];
    unsigned long long seq;             // exchanges issued so far (identical on every rank)
};

struct ny_comm {
    ny_ctx* ctx;
    int nranks, rank;
    ncclComm_t nccl;
    double* d_red;            // small device mailbox for host-value reductions
    cudaStream_t xstream;     // high-priority stream for exchanges that overlap interior kernels
    cudaEvent_t ev_ready, ev_done;
    ny_p2p p2p;
    long long n_exchanges, bytes_sent;   // face exchanges issued / bytes this rank pushed to its neighbours (ny_comm_stats)
};

// all return NY_OK or a negative ny_status; no-ops when c is null or has a single rank
int ny_comm_allreduce(ny_comm* c, double* d_buf, int n, int op_max, cudaStream_t st);
int ny_comm_allgather_inplace(ny_comm* c, double* d_recv, size_t count_per_rank, cudaStream_t st);
int ny_comm_exchange_z(ny_comm* c, double* const* arrays, int nf, size_t plane, int lo, int nint, int nh,
                       int below, int above, cudaStream_t st);
// the same for two arrays of different plane sizes / thicknesses in ONE group (a level's x and the next
// level's b after a fused down leg)
int ny_comm_exchange_z2(ny_comm* c, double* a0, size_t plane0, int nint0, double* a1, size_t plane1, int nint1, int nh,
                        int below, int above, cudaStream_t st);
