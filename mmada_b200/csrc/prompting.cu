// Sequence assembly of the generation tasks on the device: ragged pre-tokenised text + image / motion tokens ->
// (B, L) int64 ids and attention masks.
//
// Replaces the per-prompt Python list building of UniversalPrompting
//   t2i_gen_prompt  /root/reference/training/prompting_utils.py:200-233   (mode 0, task <|t2i|>, <|soi|> .. <|eoi|>)
//   t2m_prompt      :87-144 with the conditional drop-out off              (mode 0, task <|t2m|>, <|som|> .. <|eom|>)
//   mmu_gen_prompt  :379-425                                               (mode 1)
// Byte/integer work, one CTA per row; the text is read once, every output element written once (8 + 8 B per position).
#include "host_utils.h"
#include "../../include/mmada_b200.h"

namespace mmada {

struct PromptParams {
    const int64_t* text;        // all rows' text ids, concatenated
    const int64_t* text_off;    // [B + 1] offsets into `text`
    const int64_t* body;        // [B, N] image / motion tokens
    int64_t ld_body;
    int64_t* ids;               // [B, L]
    int64_t* mask;              // mode 0: [B, L] attention mask; mode 1: [B] prompt lengths
    int N, P, L, mode;
    int64_t task, bos, eos, pad, open_tok, close_tok, end_header;
};

// text with a bos in front unless it already starts with one (an empty text becomes [bos]): element i of that list
__device__ __forceinline__ int64_t text_with_bos(const int64_t* t, int len, bool add_bos, int64_t bos, int i) {
    return add_bos ? (i == 0 ? bos : t[i - 1]) : t[i];
}

__global__ void __launch_bounds__(256) build_prompts_kernel(const PromptParams p) {
    const int b = blockIdx.x;
    const int64_t* t = p.text + p.text_off[b];
    const int len = (int)(p.text_off[b + 1] - p.text_off[b]);
    const bool add_bos = len == 0 || t[0] != p.bos;
    const int tl = len + (add_bos ? 1 : 0);                         // text incl. bos
    int64_t* ids = p.ids + (int64_t)b * p.L;
    if (p.mode == 0) {
        // temp = [task] + text + [eos]; left-padded to P slots, or cut to P - 1 ids + eos
        const int n_temp = tl + 2;
        const int pad_len = p.P >= n_temp ? p.P - n_temp : 0;
        const bool cut = p.P < n_temp;
        int64_t* mask = p.mask + (int64_t)b * p.L;
        for (int i = threadIdx.x; i < p.L; i += blockDim.x) {
            int64_t v, m = 1;
            if (i < p.P) {
                if (i < pad_len) {
                    v = p.pad;
                    m = 0;
                } else {
                    const int k = i - pad_len;                      // index into temp
                    if (cut && i == p.P - 1) v = p.eos;
                    else if (k == 0) v = p.task;
                    else if (k <= tl) v = text_with_bos(t, len, add_bos, p.bos, k - 1);
                    else v = p.eos;
                }
            } else if (i == p.P) {
                v = p.open_tok;
            } else if (i < p.P + 1 + p.N) {
                v = p.body[(int64_t)b * p.ld_body + (i - p.P - 1)];
            } else {
                v = p.close_tok;
            }
            ids[i] = v;
            mask[i] = m;
        }
    } else {
        // [task] [open] body [close] | temp = text + [eos], padded with eos to P slots or cut to P - 1 ids + eos
        const int n_temp = tl + 1;
        const bool cut = p.P < n_temp;
        const int head = 3 + p.N;
        __shared__ int last_hdr;
        if (threadIdx.x == 0) last_hdr = -1;
        __syncthreads();
        for (int i = threadIdx.x; i < p.L; i += blockDim.x) {
            int64_t v;
            if (i == 0) v = p.task;
            else if (i == 1) v = p.open_tok;
            else if (i < 2 + p.N) v = p.body[(int64_t)b * p.ld_body + (i - 2)];
            else if (i == 2 + p.N) v = p.close_tok;
            else {
                const int k = i - head;                             // index into temp
                if (cut && k == p.P - 1) v = p.eos;
                else if (k < tl) v = text_with_bos(t, len, add_bos, p.bos, k);
                else v = p.eos;
                if (v == p.end_header) atomicMax(&last_hdr, k);
            }
            ids[i] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) p.mask[b] = head + (last_hdr >= 0 ? last_hdr + 1 : 0);      // prompt_length (:409-413)
    }
}

}  // namespace mmada

using namespace mmada;

extern "C" int mmada_build_prompts(const int64_t* text, const int64_t* text_off, const int64_t* body, int64_t ld_body,
                                   int64_t* ids, int64_t* mask, int B, int N, int text_slots, int mode, int64_t task_token,
                                   int64_t bos, int64_t eos, int64_t pad, int64_t open_token, int64_t close_token,
                                   int64_t end_header, void* stream) {
    if (!text_off || !body || !ids || !mask || B <= 0 || N < 0 || text_slots <= 0) return kBadArgument;
    if (mode != 0 && mode != 1) return kBadArgument;
    PromptParams p;
    p.text = text; p.text_off = text_off; p.body = body; p.ld_body = ld_body; p.ids = ids; p.mask = mask;
    p.N = N; p.P = text_slots; p.mode = mode;
    p.L = mode == 0 ? text_slots + N + 2 : 3 + N + text_slots;
    p.task = task_token; p.bos = bos; p.eos = eos; p.pad = pad; p.open_tok = open_token; p.close_tok = close_token;
    p.end_header = end_header;
    build_prompts_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(p);
    return cuda_status(cudaGetLastError());
}
