// GPU-side input preparation for the model runner: the per-token step inputs are expanded ON THE DEVICE from
// per-request state that already lives there, so the host ships O(requests) scalars per step instead of O(tokens)
// arrays and O(requests x blocks) block tables.
//
// Reference: vllm/v1/worker/gpu/input_batch.py:253-697 (Triton kernels prepare_prefill_inputs / prepare_pos_seq_lens /
// combine_sampled_and_draft_tokens), block_table.py:228-330 (_gather_block_tables / _compute_slot_mappings),
// buffer_utils.py:114-310 (staged writes).  Ours: three small sm_100a kernels over persistent device tensors
//   token_table [max_reqs, max_len] int32   every request's prompt + generated ids (sampled ids are appended on the
//                                           device by `append_sampled`, never round-tripping through the host)
//   block_rows  [groups][max_reqs, max_blocks] int32   row-ordered block tables, patched by staged diffs
// and one fused expansion kernel per step:
//   (req row, num_computed, num_scheduled) x nb  →  input_ids[T], positions[T], query_start_loc[nb+1], seq_lens[nb],
//                                                    slot_mapping[g][T], block_table[g][nb, nblk]
#include "common.cuh"

namespace {

// One CTA per scheduled request.  qsl (exclusive prefix sum of num_sched) is computed by every CTA for its own
// request with a short loop (nb <= a few thousand) — no second launch, no host-side cumsum.
__global__ void __launch_bounds__(256) prep_step_kernel(
    const int* __restrict__ rows, const int* __restrict__ num_computed, const int* __restrict__ num_sched, int nb,
    const int* __restrict__ token_table, int64_t tt_stride, const int* __restrict__ block_rows, int64_t br_stride,
    int* __restrict__ input_ids, int64_t* __restrict__ positions, int* __restrict__ qsl, int* __restrict__ seq_lens,
    int64_t* __restrict__ slot_mapping, int* __restrict__ block_table, int64_t bt_stride, int nblk, int block_size,
    const int64_t* __restrict__ pos_delta, int write_tokens, int pad_t) {
  __shared__ int s_off;
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  // exclusive prefix of num_sched[0..min(b, nb))
  int part = 0;
  for (int i = threadIdx.x; i < min(b, nb); i += blockDim.x) part += num_sched[i];
  __shared__ int red[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) red[w] = part;
  __syncthreads();
  if (w == 0) {
    int v = lane < (int)(blockDim.x >> 5) ? red[lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_off = v;
  }
  __syncthreads();
  const int off = s_off;
  if (b >= nb && pad_t > 0) {
    // mixed-batch bucket: padding REQUESTS are empty (query length 0 — the flat attention schedule skips them, their
    // "last token" index resolves to the batch's last real token), padding TOKENS [T, pad_t) sit at position 0 and
    // write to no cache slot; the extra CTA b == gridDim.x - 1 fills the token padding
    if (threadIdx.x == 0 && b < (int)gridDim.x - 1) {
      if (write_tokens) { qsl[b + 1] = off; seq_lens[b] = 1; }
      block_table[(int64_t)b * bt_stride] = 0;
    }
    if (b == (int)gridDim.x - 1)
      for (int i = off + threadIdx.x; i < pad_t; i += blockDim.x) {
        if (write_tokens) { input_ids[i] = 0; positions[i] = 0; }
        slot_mapping[i] = -1;
      }
    return;
  }
  if (b >= nb) {
    // padding slot of a decode bucket: one token at position 0 that writes to no cache slot and attends over one
    // token of the null block (same convention as the host path, worker/model_runner.py `_h2d`)
    if (threadIdx.x == 0) {
      const int idx = off + (b - nb);
      if (write_tokens) {
        input_ids[idx] = 0;
        positions[idx] = 0;
        qsl[b + 1] = idx + 1;
        seq_lens[b] = 1;
      }
      slot_mapping[idx] = -1;
      block_table[(int64_t)b * bt_stride] = 0;
    }
    return;
  }
  const int row = rows[b], nc = num_computed[b], ns = num_sched[b];
  if (threadIdx.x == 0 && write_tokens) {
    qsl[b] = off;
    if (b == nb - 1) qsl[nb] = off + ns;
    seq_lens[b] = nc + ns;
  }
  const int* toks = token_table + (int64_t)row * tt_stride;
  const int* blks = block_rows + (int64_t)row * br_stride;
  const int64_t delta = pos_delta ? pos_delta[b] : 0;       // M-RoPE: rotary position runs behind the sequence index
  for (int i = threadIdx.x; i < ns; i += blockDim.x) {
    const int pos = nc + i;
    if (write_tokens) {
      input_ids[off + i] = toks[pos];
      positions[off + i] = (int64_t)pos + delta;
    }
    slot_mapping[off + i] = (int64_t)blks[pos / block_size] * block_size + pos % block_size;
  }
  int* bt = block_table + (int64_t)b * bt_stride;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) bt[i] = blks[i];
}

// staged block-table diffs: entry e appends cnt[e] block ids (ids_flat[off[e] ..]) to row[e] starting at start[e]
__global__ void __launch_bounds__(128) block_table_diff_kernel(int* __restrict__ block_rows, int64_t br_stride,
                                                                const int* __restrict__ d_row, const int* __restrict__ d_start,
                                                                const int* __restrict__ d_cnt, const int* __restrict__ d_off,
                                                                const int* __restrict__ ids_flat) {
  pdl_trigger();
  pdl_wait();
  const int e = blockIdx.x;
  int* dst = block_rows + (int64_t)d_row[e] * br_stride + d_start[e];
  const int* src = ids_flat + d_off[e];
  for (int i = threadIdx.x; i < d_cnt[e]; i += blockDim.x) dst[i] = src[i];
}

// token_table[row[b], pos[b]] = sampled[b]  (the ids a step produced become the next step's inputs without a D2H/H2D)
__global__ void __launch_bounds__(256) append_sampled_kernel(int* __restrict__ token_table, int64_t tt_stride,
                                                              const int* __restrict__ rows, const int* __restrict__ pos,
                                                              const int64_t* __restrict__ sampled, int nb) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nb && pos[b] >= 0) token_table[(int64_t)rows[b] * tt_stride + pos[b]] = (int)sampled[b];
}

}  // namespace

EXPORT int vb_prep_step3(const int* rows, const int* num_computed, const int* num_sched, int nb, int pad_b, int pad_t,
                         const int* token_table, int64_t tt_stride, const int* block_rows, int64_t br_stride, int* input_ids,
                         int64_t* positions, int* qsl, int* seq_lens, int64_t* slot_mapping, int* block_table,
                         int64_t bt_stride, int nblk, int block_size, const int64_t* pos_delta, int write_tokens,
                         cudaStream_t st);
EXPORT int vb_prep_step2(const int* rows, const int* num_computed, const int* num_sched, int nb, int pad_b,
                         const int* token_table, int64_t tt_stride, const int* block_rows, int64_t br_stride, int* input_ids,
                         int64_t* positions, int* qsl, int* seq_lens, int64_t* slot_mapping, int* block_table,
                         int64_t bt_stride, int nblk, int block_size, const int64_t* pos_delta, int write_tokens,
                         cudaStream_t st);
EXPORT int vb_prep_step(const int* rows, const int* num_computed, const int* num_sched, int nb, const int* token_table,
                        int64_t tt_stride, const int* block_rows, int64_t br_stride, int* input_ids, int64_t* positions,
                        int* qsl, int* seq_lens, int64_t* slot_mapping, int* block_table, int64_t bt_stride, int nblk,
                        int block_size, const int64_t* pos_delta, cudaStream_t st) {
  return vb_prep_step2(rows, num_computed, num_sched, nb, nb, token_table, tt_stride, block_rows, br_stride, input_ids,
                       positions, qsl, seq_lens, slot_mapping, block_table, bt_stride, nblk, block_size, pos_delta, 1, st);
}

// pad_b >= nb: entries [nb, pad_b) are graph-bucket padding; write_tokens = 0 for the 2nd.. KV-cache group of a hybrid
// model (only slot_mapping / block_table differ between groups)
EXPORT int vb_prep_step2(const int* rows, const int* num_computed, const int* num_sched, int nb, int pad_b,
                         const int* token_table, int64_t tt_stride, const int* block_rows, int64_t br_stride, int* input_ids,
                         int64_t* positions, int* qsl, int* seq_lens, int64_t* slot_mapping, int* block_table,
                         int64_t bt_stride, int nblk, int block_size, const int64_t* pos_delta, int write_tokens,
                         cudaStream_t st) {
  return vb_prep_step3(rows, num_computed, num_sched, nb, pad_b, 0, token_table, tt_stride, block_rows, br_stride, input_ids,
                       positions, qsl, seq_lens, slot_mapping, block_table, bt_stride, nblk, block_size, pos_delta,
                       write_tokens, st);
}

// pad_t > 0: mixed-batch bucket of pad_t tokens / pad_b requests (empty padding requests + padding tokens)
EXPORT int vb_prep_step3(const int* rows, const int* num_computed, const int* num_sched, int nb, int pad_b, int pad_t,
                         const int* token_table, int64_t tt_stride, const int* block_rows, int64_t br_stride, int* input_ids,
                         int64_t* positions, int* qsl, int* seq_lens, int64_t* slot_mapping, int* block_table,
                         int64_t bt_stride, int nblk, int block_size, const int64_t* pos_delta, int write_tokens,
                         cudaStream_t st) {
  if (pad_b <= 0) return 0;
  vb_launch(prep_step_kernel, dim3(pad_t > 0 ? pad_b + 1 : pad_b), dim3(256), 0, st, rows, num_computed, num_sched, nb,
            token_table, tt_stride, block_rows, br_stride, input_ids, positions, qsl, seq_lens, slot_mapping, block_table,
            bt_stride, nblk, block_size, pos_delta, write_tokens, pad_t);
  return (int)cudaGetLastError();
}

EXPORT int vb_block_table_diff(int* block_rows, int64_t br_stride, const int* d_row, const int* d_start, const int* d_cnt,
                               const int* d_off, const int* ids_flat, int n, cudaStream_t st) {
  if (n == 0) return 0;
  vb_launch(block_table_diff_kernel, dim3(n), dim3(128), 0, st, block_rows, br_stride, d_row, d_start, d_cnt, d_off, ids_flat);
  return (int)cudaGetLastError();
}

EXPORT int vb_append_sampled(int* token_table, int64_t tt_stride, const int* rows, const int* pos, const int64_t* sampled,
                             int nb, cudaStream_t st) {
  if (nb == 0) return 0;
  vb_launch(append_sampled_kernel, dim3((nb + 255) / 256), dim3(256), 0, st, token_table, tt_stride, rows, pos, sampled, nb);
  return (int)cudaGetLastError();
}
