"""LLaDA mask predictor on the B200 kernels — host-side mirror of the reference's
``LLaDAModelLM`` (/root/reference/models/modeling_llada.py:1382-1450 -> LLaDAModel.forward :1161-1366
-> LLaDALlamaBlock.forward :886-934) for the configuration MMaDA uses: llama blocks, RMSNorm, RoPE,
SiLU-gated MLP, no biases, untied output head, no KV cache, fully bidirectional attention (the
``attention_bias`` argument is accepted and ignored exactly like the reference ignores it, Q1).

Data layout in HBM (M = batch * seq_len token rows):
  x      fp32 [M, d]        residual stream (fp32 instead of the reference's bf16: costs 2x bytes on a
                            stream read twice per layer, buys margin against the fp32 reference)
  xn     bf16 [M, d]        A operand of the q|k|v and gate/up GEMMs.  With the RMSNorm folded into the GEMMs
                            (``fused_norm``, d % 256 == 0) this is bf16(x), written by the residual GEMM's
                            epilogue together with per-tile row sums of squares ``ssq`` fp32 [M, d/256]; the
                            consumer GEMM scales its accumulator rows by rsqrt(mean(x^2)+eps) and the norm
                            weight is multiplied into the projection weight's columns at load time.
                            Otherwise: the output of the stand-alone RMSNorm kernel.
  qkv    bf16 [M, 3d]       fused q|k|v projection (weights concatenated at load time), RoPE in place
  att    bf16 [M, d]        attention output, token-major
  h      bf16 [M, ffn]      silu(ff_proj)*up_proj, produced by the GEMM epilogue (weights interleaved)
Weights are bf16, [out, in] like nn.Linear; norm weights fp32.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import ops

_P = "model.transformer."


#: reference ModelConfig options (models/configuration_llada.py:129-363) that change the arithmetic of the forward and
#: that these kernels do not implement: a config / checkpoint that sets one of them to anything but the value MMaDA-8B
#: uses is REJECTED instead of loading and silently producing other logits.  key -> accepted values
_SUPPORTED = {
    "block_type": ("llama",), "activation_type": ("silu",), "layer_norm_type": ("rms",), "rope": (True,), "alibi": (False,),
    "include_bias": (False,), "include_qkv_bias": (False, None), "bias_for_layer_norm": (False, None),
    "weight_tying": (False,), "scale_logits": (False,), "input_emb_norm": (False,), "attention_layer_norm": (False,),
    "multi_query_attention": (False, None), "layer_norm_with_affine": (True,), "rope_full_precision": (True,),
    "clip_qkv": (None,), "embedding_dropout": (0, 0.0), "residual_dropout": (0, 0.0), "attention_dropout": (0, 0.0),
}


@dataclass
class LLaDAConfig:
    """The subset of the reference's ModelConfig (models/configuration_llada.py:129-384) the path uses."""
    d_model: int = 4096
    n_heads: int = 32
    n_layers: int = 32
    mlp_hidden_size: int = 12288
    vocab_size: int = 134656
    rope_theta: float = 500000.0
    rms_norm_eps: float = 1e-5
    max_sequence_length: int = 4096
    mask_token_id: int = 126336
    embedding_size: Optional[int] = None      # rows of wte / ff_out when it differs from vocab_size (reference :309)

    @classmethod
    def from_dict(cls, d: dict) -> "LLaDAConfig":
        """Accepts the reference's full ModelConfig / HF ``config.json`` dict.  Options that would change the math
        (GQA, biases, tied head, scaled logits, q/k norms, other block / norm / activation types, ALiBi) raise."""
        d = dict(d)
        bad = []
        for k, ok in _SUPPORTED.items():
            if k in d:
                v = d[k]
                v = getattr(v, "value", v)                  # the reference's StrEnum members
                v = str(v) if isinstance(v, str) else v
                if v not in ok:
                    bad.append(f"{k}={d[k]!r} (supported: {', '.join(map(repr, ok))})")
        nh = d.get("n_heads", cls.n_heads)
        if d.get("n_kv_heads") not in (None, nh):
            bad.append(f"n_kv_heads={d['n_kv_heads']!r} != n_heads={nh} (grouped-query attention is not implemented)")
        if d.get("mlp_hidden_size") is None and "mlp_ratio" in d and "d_model" in d:
            d["mlp_hidden_size"] = int(d["mlp_ratio"]) * int(d["d_model"])
        if bad:
            raise ValueError("mmada_b200 implements the MMaDA-8B configuration of LLaDA only; unsupported: " + "; ".join(bad))
        return cls(**{k: d[k] for k in cls.__dataclass_fields__ if k in d})

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads

    @property
    def vocab_rows(self) -> int:
        return self.embedding_size or self.vocab_size


def interleave_gate_up(w_gate: torch.Tensor, w_up: torch.Tensor, block: int = 128) -> torch.Tensor:
    """[ffn, d] x2 -> [2*ffn, d] with rows alternating in blocks of 128: gate block j, up block j.
    One 256-column accumulator tile of the GEMM then holds gate and up for the same 128 outputs."""
    ffn, d = w_gate.shape
    assert ffn % block == 0, "mlp_hidden_size must be a multiple of 128"
    g = w_gate.view(ffn // block, block, d)
    u = w_up.view(ffn // block, block, d)
    return torch.stack([g, u], dim=1).reshape(2 * ffn, d).contiguous()


class CausalLMOutput:
    """Minimal stand-in for transformers' CausalLMOutputWithPast: `.logits`."""

    def __init__(self, logits: torch.Tensor):
        self.logits = logits


@dataclass
class _Layer:
    attn_norm: torch.Tensor
    wqkv: torch.Tensor
    attn_out: torch.Tensor
    ff_norm: torch.Tensor
    w_gate_up: torch.Tensor
    ff_out: torch.Tensor


class LLaDAModelLM:
    def __init__(self, config: LLaDAConfig, device="cuda", fused_norm: Optional[bool] = None):
        self.config = config
        self.device = torch.device(device)
        self.layers: List[_Layer] = []
        self.wte: Optional[torch.Tensor] = None
        self.ln_f: Optional[torch.Tensor] = None
        self.head: Optional[torch.Tensor] = None
        self._rope = None
        self.cta_group = 2
        #: RoPE in the epilogue of the fused q|k|v GEMM (needs 3*d_model % 256 == 0 and head_dim in {64, 128})
        self.fused_rope = (3 * config.d_model) % 256 == 0 and config.head_dim in (64, 128)
        #: RMSNorm folded into the GEMMs around it (see the module docstring); fixed before the weights are loaded
        can_fuse = self.fused_rope and config.d_model % 256 == 0 and config.mlp_hidden_size % 128 == 0
        if fused_norm and not can_fuse:
            raise ValueError("fused_norm needs d_model % 256 == 0, mlp_hidden_size % 128 == 0 and head_dim in {64, 128}")
        self.fused_norm = can_fuse if fused_norm is None else bool(fused_norm)
        self.masked_rows_only = True        # t2i loop: last block + head on the still-masked positions only (modeling_mmada)
        self.restrict_last_block = True     # logits_rows(rows=...) runs the last block's attn_out / MLP on those rows only
        self.kernel_launches = 0          # launches of this package's kernels (bench.py reports it)
        #: opt-in: logits_rows(rows=...) on at most this many token rows is captured once per shape into a CUDA graph and
        #: replayed (0 = off, the default).  Measured on one B200 at the 8B architecture: text-to-motion B = 1 (L = 515, ~160
        #: launches per step) 10.87 -> 10.64 ms per step, MMU B = 1 (L = 1347) 20.61 -> 20.76 ms: those steps are bound by
        #: the small-M GEMMs on the GPU, not by the host's enqueue rate, so the product leaves it off.
        self.graph_max_token_rows = 0
        self._graphs = {}

    # ---- weights ---------------------------------------------------------------------------
    def _w(self, sd, k, norm=None):
        t = sd[k].to(device=self.device)
        if norm is not None and self.fused_norm:        # fold the preceding RMSNorm's weight into the columns
            t = t.float() * sd[norm].to(device=self.device, dtype=torch.float32)[None, :]
        return t.to(dtype=torch.bfloat16).contiguous()

    def _n(self, sd, k):
        return sd[k].to(device=self.device, dtype=torch.float32).contiguous()

    def load_block(self, sd, i: int) -> None:
        """Layer ``i`` from a mapping that holds (at least) that block's tensors under the reference's key names.
        Lets a caller stream a checkpoint shard by shard / layer by layer instead of holding it whole."""
        c = self.config
        b = f"{_P}blocks.{i}."
        an, fn = b + "attn_norm.weight", b + "ff_norm.weight"
        d, f = c.d_model, c.mlp_hidden_size
        for k, shp in ((b + "q_proj.weight", (d, d)), (b + "k_proj.weight", (d, d)), (b + "v_proj.weight", (d, d)),
                       (b + "attn_out.weight", (d, d)), (b + "ff_proj.weight", (f, d)), (b + "up_proj.weight", (f, d)),
                       (b + "ff_out.weight", (d, f)), (an, (d,)), (fn, (d,))):
            if tuple(sd[k].shape) != shp:
                # e.g. a grouped-query checkpoint (k_proj / v_proj with fewer rows): refuse rather than mis-concatenate
                raise ValueError(f"{k}: shape {tuple(sd[k].shape)} does not match the configuration's {shp}")
        wqkv = torch.cat([self._w(sd, b + "q_proj.weight", an), self._w(sd, b + "k_proj.weight", an),
                          self._w(sd, b + "v_proj.weight", an)], 0).contiguous()
        layer = _Layer(self._n(sd, an), wqkv, self._w(sd, b + "attn_out.weight"), self._n(sd, fn),
                       interleave_gate_up(self._w(sd, b + "ff_proj.weight", fn), self._w(sd, b + "up_proj.weight", fn)),
                       self._w(sd, b + "ff_out.weight"))
        while len(self.layers) <= i:
            self.layers.append(None)
        self.layers[i] = layer
        self._graphs.clear()                                     # captured graphs hold the old weights' addresses

    def load_embeddings(self, sd) -> None:
        """wte, ln_f and the output head; their row count is taken from the tensors (embedding_size may exceed
        vocab_size, reference modeling_llada.py:1062,1088)."""
        self._graphs.clear()
        self.wte = self._w(sd, _P + "wte.weight")
        self.ln_f = self._n(sd, _P + "ln_f.weight")
        self.head = self._w(sd, _P + "ff_out.weight")
        d = self.config.d_model
        if self.wte.shape[1] != d or self.head.shape[1] != d or self.ln_f.numel() != d:
            raise ValueError("wte / ln_f / ff_out do not match d_model")
        if self.head.shape[0] != self.wte.shape[0]:
            raise ValueError(f"wte has {self.wte.shape[0]} rows, ff_out {self.head.shape[0]}")

    @staticmethod
    def check_state_dict_keys(keys) -> None:
        """Raise on tensors this implementation would otherwise ignore and so compute something else than the
        reference does with them: biases (include_bias / include_qkv_bias), q_norm / k_norm (attention_layer_norm),
        fused att_proj / ff_proj layouts of the non-llama block types."""
        bad = [k for k in keys
               if k.endswith(".bias") or ".q_norm." in k or ".k_norm." in k or ".att_proj." in k or ".attn_norm.bias" in k
               or k.endswith("wpe.weight") or ".emb_norm." in k]
        if bad:
            raise ValueError(f"state dict holds tensors of unsupported LLaDA options: {bad[:6]}{' ...' if len(bad) > 6 else ''}")

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> "LLaDAModelLM":
        """``sd`` uses the reference's key names.  Any Mapping works, including lazy ones that
        materialise a tensor on access (tests/test_full_size_gpu.py streams an 8B model through one)."""
        if hasattr(sd, "keys"):
            try:
                self.check_state_dict_keys(list(sd.keys()))
            except TypeError:
                pass
        self.layers = []
        self.load_embeddings(sd)
        for i in range(self.config.n_layers):
            self.load_block(sd, i)
        return self

    @classmethod
    def from_pretrained(cls, path: str, device="cuda", torch_dtype=None, fused_norm: Optional[bool] = None, **_ignored):
        """``MMadaModelLM.from_pretrained(dir, torch_dtype=torch.bfloat16)`` (reference inference_t2i.py:69): reads
        ``config.json`` (the reference's ModelConfig / MMadaConfig fields; unsupported options raise) and the (sharded)
        safetensors or ``pytorch_model.bin`` with the reference's key names, streaming one tensor at a time.  The
        kernels compute in bf16 with an fp32 residual stream whatever ``torch_dtype`` says."""
        from .checkpoint import ShardedCheckpoint, read_config
        cfg = read_config(path)
        config = cls._config_class().from_dict(cfg)
        model = cls(config, device=device, fused_norm=fused_norm)
        model.hf_config = cfg                              # codebook_size, num_vq_tokens, llm_vocab_size, ... as stored
        return model.load_state_dict(ShardedCheckpoint(path))

    @staticmethod
    def _config_class():
        return LLaDAConfig

    def init_random(self, seed: int = 0, std_scale: float = 1.0) -> "LLaDAModelLM":
        """Random weights generated on the device (benchmarks; no checkpoint is available offline).
        Scales follow the reference's 'mitchell' init (modeling_llada.py:106-110)."""
        c, dev = self.config, self.device
        self._graphs.clear()
        g = torch.Generator(device=dev).manual_seed(seed)
        d, f = c.d_model, c.mlp_hidden_size

        def rnd(shape, std):
            t = torch.empty(shape, device=dev, dtype=torch.bfloat16)
            t.normal_(0.0, std * std_scale, generator=g)
            return t

        self.wte = rnd((c.vocab_rows, d), d ** -0.5)
        self.layers = []
        for i in range(c.n_layers):
            r = (2 * (i + 1)) ** -0.5
            self.layers.append(_Layer(torch.ones(d, device=dev), rnd((3 * d, d), d ** -0.5), rnd((d, d), r * d ** -0.5),
                                      torch.ones(d, device=dev), rnd((2 * f, d), d ** -0.5), rnd((d, f), r * f ** -0.5)))
        self.ln_f = torch.ones(d, device=dev)
        self.head = rnd((c.vocab_rows, d), d ** -0.5)
        return self

    def _rope_tables(self, seq_len: int):
        # exactly the reference's table construction (modeling_llada.py:388-394), on the device
        if self._rope is None or self._rope[0].shape[0] < seq_len:
            hd = self.config.head_dim
            n = max(seq_len, 2048)
            inv_freq = 1.0 / (self.config.rope_theta ** (torch.arange(0, hd, 2, device=self.device, dtype=torch.float) / hd))
            seq = torch.arange(n, device=self.device, dtype=torch.float)
            freqs = torch.einsum("i , j -> i j", seq, inv_freq)
            self._rope = (freqs.sin().contiguous(), freqs.cos().contiguous())
        return self._rope

    # ---- forward ---------------------------------------------------------------------------
    @torch.no_grad()
    def hidden_states(self, input_ids: torch.Tensor, rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp32 residual stream after the last block (before ln_f): [B*L, d], or [len(rows), d] when ``rows`` (int32
        indices into the flattened token rows) is given — the last block then runs attn_out and the MLP on those
        rows only (after the attention every row is independent, so the values are bit-identical)."""
        c = self.config
        B, L = input_ids.shape
        sin, cos = self._rope_tables(L)
        M = B * L
        xn = torch.empty((M, c.d_model), dtype=torch.bfloat16, device=self.device)
        if self.fused_norm:
            return self._hidden_states_fused_norm(input_ids, xn, sin, cos, rows)
        x = ops.embed(input_ids.to(self.device), self.wte)
        qkv = torch.empty((M, 3 * c.d_model), dtype=torch.bfloat16, device=self.device)
        att = torch.empty((M, c.d_model), dtype=torch.bfloat16, device=self.device)
        h = torch.empty((M, c.mlp_hidden_size), dtype=torch.bfloat16, device=self.device)
        cg = self.cta_group
        last = len(self.layers) - 1
        for i, ly in enumerate(self.layers):
            ops.rmsnorm(x, ly.attn_norm, c.rms_norm_eps, out=xn)
            if self.fused_rope:
                ops.gemm_qkv_rope(xn, ly.wqkv, sin, cos, c.d_model, c.head_dim, L, out=qkv, cta_group=cg)
            else:
                ops.gemm(xn, ly.wqkv, ops.EPI_BF16, out=qkv, cta_group=cg)
                ops.rope_inplace(qkv, sin, cos, c.d_model, c.head_dim, L)
            ops.attention(qkv, B, L, c.n_heads, c.head_dim, out=att)
            if i == last and rows is not None:
                att, x = ops.gather_rows(att, rows), ops.gather_rows(x, rows)
                xn, h = xn[:rows.numel()], h[:rows.numel()]
                self.kernel_launches += 2
            ops.gemm(att, ly.attn_out, ops.EPI_RESID_F32, out=x, aux=x, cta_group=cg)
            ops.rmsnorm(x, ly.ff_norm, c.rms_norm_eps, out=xn)
            ops.gemm(xn, ly.w_gate_up, ops.EPI_SWIGLU_BF16, out=h, cta_group=cg)
            ops.gemm(h, ly.ff_out, ops.EPI_RESID_F32, out=x, aux=x, cta_group=cg)
        self.kernel_launches += 1 + (7 if self.fused_rope else 8) * len(self.layers)
        return x

    def _hidden_states_fused_norm(self, input_ids, xn, sin, cos, rows=None):
        """The block stack with every attn_norm / ff_norm folded into the GEMMs around it: 5 launches per layer
        (q|k|v+RoPE, attention, attn_out+residual, gate/up+SwiGLU, ff_out+residual)."""
        c = self.config
        B, L = input_ids.shape
        M, d, eps, cg = B * L, c.d_model, c.rms_norm_eps, self.cta_group
        tiles = d // 256
        ssq = torch.empty((M, tiles), dtype=torch.float32, device=self.device)
        qkv = torch.empty((M, 3 * d), dtype=torch.bfloat16, device=self.device)
        att = torch.empty((M, d), dtype=torch.bfloat16, device=self.device)
        h = torch.empty((M, c.mlp_hidden_size), dtype=torch.bfloat16, device=self.device)
        x = ops.embed_norm(input_ids.to(self.device), self.wte, xn, ssq)
        t_in = 1                                                # the embedding writes one partial per row
        last = len(self.layers) - 1
        for i, ly in enumerate(self.layers):
            ops.gemm_qkv_rope_rownorm(xn, ly.wqkv, sin, cos, d, c.head_dim, L, ssq, t_in, d, eps, out=qkv, cta_group=cg)
            ops.attention(qkv, B, L, c.n_heads, c.head_dim, out=att)
            if i == last and rows is not None:                  # the caller reads these rows only
                att, x = ops.gather_rows(att, rows), ops.gather_rows(x, rows)
                xn, h, ssq = xn[:rows.numel()], h[:rows.numel()], ssq[:rows.numel()]
                self.kernel_launches += 2
            ops.gemm_resid_norm(att, ly.attn_out, x, xn, ssq, cta_group=cg)
            ops.gemm_swiglu_rownorm(xn, ly.w_gate_up, ssq, tiles, d, eps, out=h, cta_group=cg)
            if i == last:                                       # ln_f reads the fp32 stream (row-gathered)
                ops.gemm(h, ly.ff_out, ops.EPI_RESID_F32, out=x, aux=x, cta_group=cg)
            else:
                ops.gemm_resid_norm(h, ly.ff_out, x, xn, ssq, cta_group=cg)
            t_in = tiles
        self.kernel_launches += 1 + 5 * len(self.layers)
        return x

    @torch.no_grad()
    def logits_rows(self, input_ids: torch.Tensor, rows: Optional[torch.Tensor], col_lo: int = 0,
                    col_hi: Optional[int] = None) -> torch.Tensor:
        """fp32 logits for the token rows ``rows`` (int32 indices into the flattened [B*L] rows; None = all)
        and vocabulary columns [col_lo, col_hi): ln_f and the output head run on those rows only."""
        c = self.config
        if (rows is not None and 0 < input_ids.numel() <= self.graph_max_token_rows and not torch.cuda.is_current_stream_capturing()
                and input_ids.is_cuda):
            return self._graphed_logits_rows(input_ids, rows, col_lo, col_hi)
        return self._logits_rows(input_ids, rows, col_lo, col_hi)

    def _logits_rows(self, input_ids, rows, col_lo=0, col_hi=None):
        c = self.config
        # rows that are a small enough share of the batch also restrict the last block (0.8 % of a config-2 step)
        early = rows is not None and self.restrict_last_block and 10 * rows.numel() <= 9 * input_ids.numel()
        x = self.hidden_states(input_ids, rows if early else None)
        xn = ops.rmsnorm(x, self.ln_f, c.rms_norm_eps, rows=None if early else rows)
        col_hi = self.head.shape[0] if col_hi is None else col_hi
        self.kernel_launches += 2
        return ops.gemm(xn, self.head[col_lo:col_hi], ops.EPI_F32, cta_group=self.cta_group)

    def _graphed_logits_rows(self, input_ids, rows, col_lo, col_hi):
        """One CUDA graph per (batch, length, number of rows, column range): captured on first use after an eager warm-up
        call (which also settles one-time state: RoPE tables, kernel attributes), replayed afterwards with the ids and
        the row list copied into the graph's static inputs.  Same kernels, same order, same results."""
        key = (tuple(input_ids.shape), rows.numel(), col_lo, col_hi, self.restrict_last_block, self.fused_norm,
               self.fused_rope, self.cta_group)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 16:                          # bounded: every graph owns its activations
                self._graphs.pop(next(iter(self._graphs)))
            ids_buf, rows_buf = input_ids.clone(), rows.clone()
            n0 = self.kernel_launches
            with torch.cuda.device(self.device):
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    self._logits_rows(ids_buf, rows_buf, col_lo, col_hi)
                torch.cuda.current_stream(self.device).wait_stream(side)
                n_launches = self.kernel_launches - n0
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self._logits_rows(ids_buf, rows_buf, col_lo, col_hi)
            self.kernel_launches = n0
            g = self._graphs[key] = (graph, ids_buf, rows_buf, out, n_launches)
        graph, ids_buf, rows_buf, out, n_launches = g
        ids_buf.copy_(input_ids)
        rows_buf.copy_(rows)
        graph.replay()
        self.kernel_launches += n_launches
        return out.clone()                                       # the graph's output buffer is overwritten by the next replay

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_bias=None, **_ignored) -> CausalLMOutput:
        """Drop-in ``model(input_ids, attention_bias=...).logits`` -> (B, L, V) fp32.  ``attention_bias``
        is ignored, as in the reference (modeling_llada.py:711-718 passes attn_mask=None)."""
        B, L = input_ids.shape
        lg = self.logits_rows(input_ids, None)
        return CausalLMOutput(lg.view(B, L, -1))

    __call__ = forward
