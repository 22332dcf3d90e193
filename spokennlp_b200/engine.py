"""Host-side orchestration of the encoder hot path: parameter packing and the per-layer
kernel schedule (forward and backward) over the C ABI.  No arithmetic happens in Python.

Data layout in HBM:
  * one flat fp32 buffer holds every parameter (HF order below); the module's nn.Parameters are VIEWS into it, so
    checkpoints load/save with HF keys while query/key/value sit adjacent and form the packed [3H,H] QKV weight;
  * one flat fp16 buffer mirrors it (tensor-core operands); one flat fp32 buffer holds the gradients;
  * activations are token-major [B*S, width] fp16; Q/K/V live packed in one [B*S, 3H] buffer and are addressed by
    TMA coordinates (no head split / permute); pre-LayerNorm sums are kept in fp32.

Reference schedule being replaced: BertLayer.forward (mmvts/src/models/cross_encoder/bert_model.py:518-553) and its
autograd; embeddings (:184-210).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .lib import B200Error
from .blocks import AttnSaved, AttnWeights, FfnSaved, FfnWeights, attn_block_bwd, attn_block_fwd, ffn_block_bwd, ffn_block_fwd

Tensor = torch.Tensor
ALIGN = 64  # elements


def layer_param_names(i: int) -> List[str]:
    p = f"encoder.layer.{i}."
    return [p + "attention.self.query.weight", p + "attention.self.key.weight", p + "attention.self.value.weight",
            p + "attention.self.query.bias", p + "attention.self.key.bias", p + "attention.self.value.bias",
            p + "attention.output.dense.weight", p + "attention.output.dense.bias",
            p + "attention.output.LayerNorm.weight", p + "attention.output.LayerNorm.bias",
            p + "intermediate.dense.weight", p + "intermediate.dense.bias",
            p + "output.dense.weight", p + "output.dense.bias",
            p + "output.LayerNorm.weight", p + "output.LayerNorm.bias"]


EMB_NAMES = ["embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
             "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"]


class FlatParams:
    """Packs named fp32 parameters, in the given order, into one flat buffer (+ fp16 mirror, + optional grad buffer)
    and re-points the nn.Parameters at views of it."""

    def __init__(self, named: Sequence[Tuple[str, torch.nn.Parameter]], device):
        self.names = [n for n, _ in named]
        self.params = {n: p for n, p in named}
        self.offsets: Dict[str, int] = {}
        off = 0
        for n in self.names:                      # every offset is 64-element aligned; tensors whose size is a
            self.offsets[n] = off                 # multiple of 64 (all H x H / H-sized ones) therefore stay adjacent
            off = (off + self.params[n].numel() + ALIGN - 1) // ALIGN * ALIGN
        self.numel = off
        self.flat32 = torch.zeros(off, dtype=torch.float32, device=device)
        self.flat16 = torch.empty(off, dtype=torch.float16, device=device)
        self.grad32: Optional[Tensor] = None
        with torch.no_grad():
            for n, p in self.params.items():
                v = self.view32(n)
                v.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = v
        self.signature = self._sig()
        self.version = -1
        self.sync_half()

    def _sig(self):
        return tuple(p.data_ptr() for p in self.params.values())

    def intact(self, named: Optional[Sequence[Tuple[str, torch.nn.Parameter]]] = None) -> bool:
        """True while the packed buffers still ARE the module's parameters.  `named` = the module's CURRENT (name, Parameter)
        list: a Parameter object that was replaced (older transformers' `resize_token_embeddings` -> `set_input_embeddings`,
        `load_state_dict(assign=True)`, a swapped sub-module — ts_sentence_seq_labeling.py:284 resizes the vocabulary) leaves the
        OLD object pointing into the flat buffer, so comparing only the packed objects' own data_ptrs would call a stale,
        smaller table intact; the identity of every current parameter is compared as well."""
        if self._sig() != self.signature:
            return False
        if named is None:
            return True
        if len(named) != len(self.names):
            return False
        for n, p in named:
            q = self.params.get(n)
            if q is not p or p.data_ptr() != self.flat32.data_ptr() + 4 * self.offsets[n] or p.shape != q.shape:
                return False
        return True

    def _view(self, flat: Tensor, name: str, n_extra: Sequence[str] = ()) -> Tensor:
        p = self.params[name]
        o = self.offsets[name]
        if not n_extra:
            return flat[o:o + p.numel()].view(p.shape)
        rows = p.shape[0] + sum(self.params[e].shape[0] for e in n_extra)
        numel = p.numel() + sum(self.params[e].numel() for e in n_extra)
        end = o
        for e in (name,) + tuple(n_extra):        # the packed view is only valid if the members are adjacent
            assert self.offsets[e] == end, f"{e} is not adjacent in the flat buffer"
            end += self.params[e].numel()
        return flat[o:o + numel].view(rows, *p.shape[1:])

    def view32(self, name, extra=()):
        return self._view(self.flat32, name, extra)

    def view16(self, name, extra=()):
        return self._view(self.flat16, name, extra)

    def viewg(self, name, extra=()):
        return self._view(self.grad32, name, extra)

    def cur_version(self) -> int:
        return sum(p._version for p in self.params.values())

    def sync_half(self, force: bool = False) -> None:
        """Refresh the fp16 mirror if any parameter was modified in place since the last refresh.  Under stream capture the refresh is
        always issued, so that it becomes part of the graph: a replay after an optimizer step must see the new weights, and the version
        check only runs at capture time (spokennlp_b200.graphs.GraphedStep)."""
        v = self.cur_version()
        if force or v != self.version or (self.flat32.is_cuda and torch.cuda.is_current_stream_capturing()):
            ops.cast_f32_to_f16(self.flat32, self.flat16)
            self.version = v

    def ensure_grad(self, zero: bool = True) -> Tensor:
        if self.grad32 is None:
            self.grad32 = torch.zeros_like(self.flat32)
        elif zero:
            self.grad32.zero_()
        return self.grad32


def next_drop_seed(owner, device) -> Tensor:
    """The base dropout seed of one forward of `owner` (a drop-in module): an int32 [1] device tensor that stays untouched until that
    forward's backward has regenerated its masks.  The module's running seed lives on the device and advances there (one add per
    forward; this call's copy is handed out), so a forward captured into a CUDA graph draws fresh masks on every replay; its first
    value comes from torch's CPU generator, so `torch.manual_seed` governs it as it governs `nn.Dropout` in the reference."""
    base = owner.__dict__.get("_b200_seed_dev")
    if base is None or base.device != torch.device(device):
        if torch.cuda.is_current_stream_capturing():
            raise B200Error("run the step once before capturing it: the dropout seed is created on first use")
        base = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int32).to(device)
        object.__setattr__(owner, "_b200_seed_dev", base)
    else:
        base.add_(0x632BE5AB)                       # (int32 wrap-around is fine: the kernels hash seed, site and element index)
    return base.clone()


@dataclass
class LayerViews:
    attn: AttnWeights
    ffn: FfnWeights


@dataclass
class LayerSaved:
    attn: AttnSaved = None
    ffn: FfnSaved = None


@dataclass
class Saved:
    B: int = 0; S: int = 0
    ids: Optional[Tensor] = None; tt: Optional[Tensor] = None; pos: Optional[Tensor] = None
    inputs_embeds: Optional[Tensor] = None
    key_bias: Optional[Tensor] = None; kv_len: Optional[Tensor] = None
    pack: Optional[ops.RowIndex] = None
    layers: List[LayerSaved] = field(default_factory=list)
    out: Tensor = None
    drop_emb: Optional[ops.Dropout] = None


class DropPlan:
    """Where and how hard dropout applies in one training forward (HF BertConfig.hidden_dropout_prob /
    attention_probs_dropout_prob).  `seed` is a 1-element int32 DEVICE tensor holding this step's base seed; each site
    hashes it with its own id, the backward regenerates the masks from the same (seed, site)."""
    EMB, HEAD = 0xE000, 0xE001
    ATTN, ATTN_OUT, FFN_OUT, XATTN, XATTN_OUT = range(5)

    def __init__(self, seed: Tensor, p_hidden: float, p_attn: float):
        self.seed, self.p_hidden, self.p_attn = seed, float(p_hidden), float(p_attn)

    def at(self, site: int, p: float) -> Optional[ops.Dropout]:
        return ops.Dropout(self.seed, site, p) if p > 0.0 else None

    def layer(self, i: int, kind: int) -> Optional[ops.Dropout]:
        return self.at(i * 8 + kind, self.p_attn if kind in (self.ATTN, self.XATTN) else self.p_hidden)


class EncoderEngine:
    """Runs embeddings + L encoder layers on the packed parameters."""

    def __init__(self, flat: FlatParams, hidden: int, heads: int, inter: int, n_layers: int, eps: float,
                 prefix_layers: str = "encoder.layer.", pad_id: Optional[int] = None):
        """`pad_id` = config.pad_token_id: nn.Embedding(padding_idx=...) of BertEmbeddings (bert_model.py:171) — that row of
        the word table takes part in the forward and receives no gradient."""
        assert hidden == heads * 64, "the sm_100a attention kernels are specialised for head_dim 64"
        self.flat, self.H, self.heads, self.I, self.L, self.eps = flat, hidden, heads, inter, n_layers, eps
        self.prefix = prefix_layers
        self.pad_id = pad_id

    # ---- views -------------------------------------------------------------------------------------------------
    def _lv(self, i: int, kind: str) -> LayerViews:
        n = layer_param_names(i)
        f = self.flat
        w = {"p": f.view16, "g": f.viewg}[kind]
        s = {"p": f.view32, "g": f.viewg}[kind]
        return LayerViews(attn=AttnWeights(wqkv=w(n[0], (n[1], n[2])), bqkv=s(n[3], (n[4], n[5])), wo=w(n[6]), bo=s(n[7]),
                                           g=s(n[8]), b=s(n[9])),
                          ffn=FfnWeights(w1=w(n[10]), bf1=s(n[11]), w2=w(n[12]), bf2=s(n[13]), g=s(n[14]), b=s(n[15])))

    def layer(self, i):
        c = self.__dict__.setdefault("_pcache", {})
        key = (i, self.flat.flat16.data_ptr())
        if key not in c:
            c[key] = self._lv(i, "p")
        return c[key]

    def layer_grads(self, i):
        c = self.__dict__.setdefault("_gcache", {})
        key = (i, self.flat.grad32.data_ptr())
        if key not in c:
            if len(c) > 4 * self.L:
                c.clear()
            c[key] = self._lv(i, "g")
        return c[key]

    # ---- forward -----------------------------------------------------------------------------------------------
    def embed(self, ids, tt, pos, inputs_embeds, B, S, drop=None, rows: Optional[int] = None):
        """Returns the embedding output twice: fp16 (tensor-core operand) and fp32 (residual stream).  `rows`: packed row count."""
        f = self.flat
        M = B * S if rows is None else rows
        y32 = torch.empty(M, self.H, dtype=torch.float32, device=f.flat32.device)
        y16 = ops.embed_ln_fwd(ids, tt, pos, inputs_embeds, f.view32(EMB_NAMES[0]), f.view32(EMB_NAMES[1]),
                               f.view32(EMB_NAMES[2]), f.view32(EMB_NAMES[3]), f.view32(EMB_NAMES[4]), self.eps, M, S,
                               self.H, y32=y32, drop=drop)
        return y16, y32

    def layer_forward(self, p: LayerViews, x: Tensor, x32: Tensor, B: int, S: int, key_bias, kv_len, save: bool,
                      want_probs: bool = False, drop: Optional[DropPlan] = None, index: int = 0, owns_input: bool = False,
                      pack: Optional[ops.RowIndex] = None):
        """One BertLayer (bert_model.py:518-553).  x: fp16 layer input (GEMM operand), x32: the same activations in fp32
        (residual stream: keeping the skip connection un-rounded holds the 12-layer hidden-state error under 1e-3)."""
        d = (lambda k: drop.layer(index, k)) if drop is not None else (lambda k: None)
        a16, a32, sva, probs = attn_block_fwd(p.attn, x, x32, B, S, self.heads, self.eps, key_bias, kv_len, save=save,
                                              want_probs=want_probs, drop_attn=d(DropPlan.ATTN), drop_hidden=d(DropPlan.ATTN_OUT),
                                              owns_residual=owns_input, pack=pack)
        # a32 (the attention block's fp32 output) is never handed out, so the FFN block always owns its residual
        y16, y32, svf = ffn_block_fwd(p.ffn, a16, a32, self.eps, save=save, drop_hidden=d(DropPlan.FFN_OUT), owns_residual=True)
        return y16, y32, (LayerSaved(attn=sva, ffn=svf) if save else None), probs

    def forward(self, ids, tt, pos, inputs_embeds, key_bias, kv_len, B: int, S: int, *, save: bool,
                want_hidden: bool = False, want_probs: bool = False, drop: Optional[DropPlan] = None,
                pack: Optional[ops.RowIndex] = None):
        """`drop` (training only) turns on the reference's dropout sites; None = eval / p=0.  `pack` (ops.compact_rows of the
        attention mask): ids / tt / pos are the PACKED valid tokens (pos = position inside the sequence) and every kernel runs on
        pack.n rows instead of B*S; key_bias / kv_len must be None (the sequence lengths ARE the mask)."""
        self.flat.sync_half()
        if pack is not None and (key_bias is not None or kv_len is not None or inputs_embeds is not None or pos is None):
            raise ops.L.B200Error("packed rows: pass packed ids / token types / positions and no key mask")
        drop_emb = drop.at(DropPlan.EMB, drop.p_hidden) if drop is not None else None
        x, x32 = self.embed(ids, tt, pos, inputs_embeds, B, S, drop=drop_emb, rows=None if pack is None else pack.n)
        saved = Saved(B=B, S=S, ids=ids, tt=tt, pos=pos, inputs_embeds=inputs_embeds if ids is None else None, key_bias=key_bias,
                      kv_len=kv_len, pack=pack, drop_emb=drop_emb) if save else None
        hiddens, probs_all = ([x32] if want_hidden else None), ([] if want_probs else None)
        for i in range(self.L):
            # a layer input that is also returned as a hidden state must survive the layer (blocks.Experimental.resadd)
            x, x32, sv, probs = self.layer_forward(self.layer(i), x, x32, B, S, key_bias, kv_len, save, want_probs, drop, i,
                                                   owns_input=not want_hidden, pack=pack)
            if save:
                saved.layers.append(sv)
            if want_hidden:
                hiddens.append(x32)
            if want_probs:
                probs_all.append(probs)
        if save:
            saved.out = x
        return x, x32, saved, hiddens, probs_all

    # ---- backward ----------------------------------------------------------------------------------------------
    def layer_backward(self, p: LayerViews, g: LayerViews, sv: LayerSaved, dy: Tensor, B: int, S: int, key_bias, kv_len,
                       inv_scale: Optional[Tensor], ws: Tensor) -> Tensor:
        """dy: fp16 gradient wrt the layer output (scaled by the loss scale).  Weight/bias/LN gradients are accumulated
        (+=) into the fp32 views `g`, multiplied by *inv_scale.  Returns the gradient wrt the layer input."""
        d_attn_out = ffn_block_bwd(p.ffn, g.ffn, sv.ffn, dy, inv_scale)
        dx, _ = attn_block_bwd(p.attn, g.attn, sv.attn, d_attn_out, B, S, self.heads, key_bias, kv_len, inv_scale, ws)
        return dx

    def backward(self, saved: Saved, dy: Tensor, inv_scale: Optional[Tensor], *, embeddings: bool = True,
                 after_layer=None, want_d_inputs_embeds: bool = False) -> Optional[Tensor]:
        """Accumulates all parameter gradients into flat.grad32 (must exist).  `after_layer(i)` is called once layer
        i's gradients are complete (the data-parallel trainer launches that layer's allreduce there).  When the forward ran
        on `inputs_embeds`, position / type / LayerNorm gradients are still produced and the gradient wrt `inputs_embeds`
        (fp32 [B*S, H]) is returned if asked for."""
        B, S = saved.B, saved.S
        f = self.flat
        rows = B * S if saved.pack is None else saved.pack.n
        ws = ops.attn_bwd_workspace(B, self.heads, S, dy.device, rows=None if saved.pack is None else rows)
        for i in reversed(range(self.L)):
            dy = self.layer_backward(self.layer(i), self.layer_grads(i), saved.layers[i], dy, B, S, saved.key_bias,
                                     saved.kv_len, inv_scale, ws)
            saved.layers[i] = None          # release this layer's activations
            if after_layer is not None:
                after_layer(i)
        d_emb = None
        if embeddings and (saved.ids is not None or saved.inputs_embeds is not None):
            if saved.ids is None and want_d_inputs_embeds:
                d_emb = torch.empty(rows, self.H, dtype=torch.float32, device=dy.device)
            ops.embed_ln_bwd(dy, None, saved.ids, saved.tt, saved.pos, f.view32(EMB_NAMES[0]), f.view32(EMB_NAMES[1]),
                             f.view32(EMB_NAMES[2]), f.view32(EMB_NAMES[3]), f.viewg(EMB_NAMES[0]), f.viewg(EMB_NAMES[1]),
                             f.viewg(EMB_NAMES[2]), f.viewg(EMB_NAMES[3]), f.viewg(EMB_NAMES[4]), inv_scale, self.eps, rows,
                             S, self.H, drop=saved.drop_emb, pad_id=self.pad_id, inputs_embeds=saved.inputs_embeds,
                             d_inputs_embeds=d_emb)
        if after_layer is not None:
            after_layer(-1)
        return d_emb
