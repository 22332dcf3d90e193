"""Drop-in `PoNetModel` for alimeeting4mug's topic-segmentation script, backed by libb200enc.so.

Boundary: `modelscope.models.nlp.ponet.PoNetModel` as constructed and called at
alimeeting4mug/src/models/modeling_ponet.py:41 (`PoNetModel(config, add_pooling_layer=False)`) and :68-79
(`self.ponet(input_ids, attention_mask=..., token_type_ids=..., segment_ids=..., position_ids=..., head_mask=...,
inputs_embeds=..., output_attentions=..., output_hidden_states=..., return_dict=...)`).  modelscope is not installable
here, so the layer arithmetic follows the restatement in oracle/ponet_oracle.py (**parity unpinned**) and
the parameter names follow the public PoNet implementation (`attention.self.dense_{q,k,o,segment,local}`).

Per layer: ONE packed [5H,H] tcgen05 GEMM produces Q|K|O|Sg|Lc; the pooling mixer (global softmax-pooled vector per
head, segment max, local max-3, fuse) runs as coalesced HBM-bound kernels (`b200_ponet_mix_fwd`); the output dense +
residual + LayerNorm and the feed-forward block are the same kernels as BERT.  This round ships the forward
(inference / predict path); the mixer backward is listed as next in DESIGN.md §7.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn
from transformers import PretrainedConfig
from transformers.modeling_outputs import BaseModelOutputWithPoolingAndCrossAttentions

from . import ops
from .blocks import FfnWeights, ffn_block_fwd
from .engine import EMB_NAMES, FlatParams
from .lib import B200Error
from .modeling_bert import BertEmbeddings, BertIntermediate, BertOutput, BertPooler, BertSelfOutput

F16, F32 = torch.float16, torch.float32
_PROJ = ("q", "k", "o", "segment", "local")


class PoNetConfig(PretrainedConfig):
    model_type = "ponet"

    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=512,
                 type_vocab_size=2, initializer_range=0.02, layer_norm_eps=1e-12, pad_token_id=0, **kwargs):
        super().__init__(pad_token_id=pad_token_id, **kwargs)
        self.vocab_size, self.hidden_size, self.num_hidden_layers = vocab_size, hidden_size, num_hidden_layers
        self.num_attention_heads, self.intermediate_size, self.hidden_act = num_attention_heads, intermediate_size, hidden_act
        self.hidden_dropout_prob, self.attention_probs_dropout_prob = hidden_dropout_prob, attention_probs_dropout_prob
        self.max_position_embeddings, self.type_vocab_size = max_position_embeddings, type_vocab_size
        self.initializer_range, self.layer_norm_eps = initializer_range, layer_norm_eps


class PoNetSelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        for n in _PROJ:
            setattr(self, f"dense_{n}", nn.Linear(config.hidden_size, config.hidden_size))


class PoNetAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = PoNetSelfAttention(config)
        self.output = BertSelfOutput(config)


class PoNetLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = PoNetAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)


class PoNetEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([PoNetLayer(config) for _ in range(config.num_hidden_layers)])


def ponet_layer_names(i: int) -> List[str]:
    p = f"encoder.layer.{i}."
    return ([p + f"attention.self.dense_{n}.weight" for n in _PROJ] + [p + f"attention.self.dense_{n}.bias" for n in _PROJ] +
            [p + "attention.output.dense.weight", p + "attention.output.dense.bias", p + "attention.output.LayerNorm.weight",
             p + "attention.output.LayerNorm.bias", p + "intermediate.dense.weight", p + "intermediate.dense.bias",
             p + "output.dense.weight", p + "output.dense.bias", p + "output.LayerNorm.weight", p + "output.LayerNorm.bias"])


class PoNetModel(nn.Module):
    def __init__(self, config, add_pooling_layer: bool = True):
        super().__init__()
        self.config = config
        if config.hidden_size != 64 * config.num_attention_heads:
            raise B200Error("B200 PoNetModel requires head_dim == 64")
        self.embeddings = BertEmbeddings(config)
        self.encoder = PoNetEncoder(config)
        self.pooler = BertPooler(config) if add_pooling_layer else None
        std = getattr(config, "initializer_range", 0.02)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data.normal_(mean=0.0, std=std)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()
        self._flat: Optional[FlatParams] = None

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def set_input_embeddings(self, value):
        self.embeddings.word_embeddings = value

    def _packed(self, device) -> FlatParams:
        f = self._flat
        if f is not None and f.intact():
            f.sync_half()
            return f
        if torch.device(device).type != "cuda":
            raise B200Error("B200 PoNetModel runs on CUDA devices only (no CPU fallback)")
        own = dict(self.named_parameters())
        names = list(EMB_NAMES)
        for i in range(self.config.num_hidden_layers):
            names += ponet_layer_names(i)
        self._flat = FlatParams([(n, own[n]) for n in names], device)
        return self._flat

    @torch.no_grad()
    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, segment_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, output_attentions=None, output_hidden_states=None, return_dict=None, **kwargs):
        cfg = self.config
        if head_mask is not None or output_attentions:
            raise B200Error("head_mask / output_attentions are not supported by the B200 PoNetModel")
        if input_ids is None:
            raise B200Error("B200 PoNetModel needs input_ids")
        if not input_ids.is_cuda:
            raise B200Error("B200 PoNetModel got CPU tensors: there is no CPU fallback")
        if segment_ids is None:
            raise B200Error("PoNetModel.forward needs segment_ids (modeling_ponet.py:72)")
        return_dict = True if return_dict is None else return_dict
        B, S = input_ids.shape
        f = self._packed(input_ids.device)
        H, heads, eps, dev = cfg.hidden_size, cfg.num_attention_heads, float(cfg.layer_norm_eps), input_ids.device
        M = B * S
        if position_ids is None:
            position_ids = self.embeddings.position_ids[:, :S].expand(B, S)     # honours the driver's in-place 4096 tiling
        pos = position_ids.expand(B, S).contiguous().view(-1)
        tt = token_type_ids.contiguous().view(-1) if token_type_ids is not None else None
        key_bias = None
        if attention_mask is not None:
            key_bias, _ = ops.mask_to_bias(attention_mask)
        seg = segment_ids.contiguous().to(torch.int64)
        nseg = S + 2
        x32 = torch.empty(M, H, dtype=F32, device=dev)
        x16 = ops.embed_ln_fwd(input_ids.contiguous().view(-1), tt, pos, None, f.view32(EMB_NAMES[0]), f.view32(EMB_NAMES[1]),
                               f.view32(EMB_NAMES[2]), f.view32(EMB_NAMES[3]), f.view32(EMB_NAMES[4]), eps, M, S, H, y32=x32)
        hiddens = [x32.view(B, S, H)] if output_hidden_states else None
        for i in range(cfg.num_hidden_layers):
            n = ponet_layer_names(i)
            proj = torch.empty(M, 5 * H, dtype=F16, device=dev)
            ops.gemm(x16, f.view16(n[0], tuple(n[1:5])), proj, epilogue=ops.EPI_BIAS, bias=f.view32(n[5], tuple(n[6:10])))
            mix = torch.empty(M, H, dtype=F16, device=dev)
            ops.ponet_mix_fwd(proj, seg, mix, B, S, heads, nseg, key_bias=key_bias)
            pre = torch.empty(M, H, dtype=F32, device=dev)
            ops.gemm(mix, f.view16(n[10]), pre, epilogue=ops.EPI_BIAS_RES32, bias=f.view32(n[11]), aux=x32)
            a32 = torch.empty(M, H, dtype=F32, device=dev)
            a16 = ops.layernorm_fwd(pre, f.view32(n[12]), f.view32(n[13]), eps, y32=a32)
            ffn = FfnWeights(w1=f.view16(n[14]), bf1=f.view32(n[15]), w2=f.view16(n[16]), bf2=f.view32(n[17]), g=f.view32(n[18]),
                             b=f.view32(n[19]))
            x16, x32, _ = ffn_block_fwd(ffn, a16, a32, eps, save=False)
            if output_hidden_states:
                hiddens.append(x32.view(B, S, H))
        seq = x32.view(B, S, H)
        pooled = self.pooler(seq) if self.pooler is not None else None
        hs = tuple(hiddens) if output_hidden_states else None
        if not return_dict:
            return (seq, pooled) + ((hs,) if hs is not None else ())
        return BaseModelOutputWithPoolingAndCrossAttentions(last_hidden_state=seq, pooler_output=pooled, hidden_states=hs, attentions=None)
