// pixelrec_b200 -- A1: on-device construction of SEQTrainDataset batches (REC/data/dataset/trainset.py:40-75).
//   The reference builds every sample in Python (10 DataLoader workers: rejection-sampled negatives, left padding,
//   mask) -- ~1e5 sequences/s at best, far below what the GPU step consumes (4.6e5 seq/s on one B200).  Here the padded
//   training windows live in HBM and one launch assembles the step's tensors:
//     items[b,0,:] = padded[sel[b],:]                         positives, left-padded with 0        (trainset.py:46-50)
//     items[b,1,t] = uniform draw from [1, item_num-1] that is not one of the sequence's items,
//                    for every t after the first real item; 0 elsewhere                            (trainset.py:40-44,52-63)
//     mask[b,t-1]  = 1 for those t, else 0
//   Random numbers: Philox4x32-10(seed; counter = (b*W + t)*64 + attempt, stream 0x5eed), first word, mapped with a
//   multiply-high (no modulo bias); oracle/sasrec_np.py seq_batch_build restates it, so the output is bit-exact.
#include "common.cuh"

namespace pr {

__global__ void __launch_bounds__(256) seq_batch_kernel(const long long* __restrict__ padded, long long n_seq, int W,
                                                        const long long* __restrict__ sel, long long B, long long item_num,
                                                        unsigned long long seed, long long* __restrict__ items,
                                                        long long* __restrict__ mask, int* __restrict__ status) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * W) return;
    const long long b = i / W;
    const int t = (int)(i - b * W);
    long long s = sel[b];
    if (s < 0 || s >= n_seq) {
        if (status) atomicOr(status, 1);
        s = 0;
    }
    const long long* row = padded + s * W;
    const long long pos = row[t];
    int first = W;                                   // index of the first real item (left padding before it)
    for (int k = 0; k < W; ++k)
        if (row[k] != 0) { first = k; break; }
    long long neg = 0;
    if (t > first) {
        const Philox ph(seed);
        const unsigned long long range = (unsigned long long)(item_num - 1);
        for (int attempt = 0; attempt < 64; ++attempt) {
            const uint4 r = ph((unsigned long long)(b * W + t) * 64ull + attempt, 0x5eedu);
            neg = 1 + (long long)(((unsigned long long)r.x * range) >> 32);
            bool clash = false;
            for (int k = first; k < W; ++k) clash |= (row[k] == neg);
            if (!clash) break;
        }
    }
    items[(b * 2 + 0) * W + t] = pos;
    items[(b * 2 + 1) * W + t] = neg;
    if (t >= 1) mask[b * (W - 1) + (t - 1)] = (t > first) ? 1 : 0;
}

}  // namespace pr

using namespace pr;

extern "C" int pr_seq_batch_build(const int64_t* padded, int64_t n_seq, int W, const int64_t* sel, int64_t B, int64_t item_num,
                                  uint64_t seed, int64_t* items, int64_t* mask, int32_t* status, pr_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PR_CHECK_ARG(n_seq > 0 && W >= 2 && W <= 1024 && B >= 0 && item_num >= 2, "pr_seq_batch_build: bad shape n_seq=%lld W=%d B=%lld item_num=%lld",
                 (long long)n_seq, W, (long long)B, (long long)item_num);
    if (B == 0) return PR_OK;
    PR_CHECK_ARG(padded && sel && items && mask, "pr_seq_batch_build: null pointer");
    const long long n = B * W;
    seq_batch_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>((const long long*)padded, n_seq, W, (const long long*)sel, B,
                                                                 item_num, seed, (long long*)items, (long long*)mask, status);
    PR_CUDA_LAUNCH_CHECK("seq_batch_kernel");
    return PR_OK;
}
