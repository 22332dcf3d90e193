"""spokennlp_b200 — B200-native (sm_100a) transformer-encoder hot path for SpokenNLP's scripts.

Public surface:
  * `BertModel`            drop-in for transformers' BertModel (same signature / state_dict), CUDA library inside
  * `patch_transformers()` make the reference's own wrappers pick it up
  * `ops`                  tensor-level wrappers over the C ABI (include/b200enc.h)
  * `lib`                  ctypes loader / builder of libb200enc.so
"""
from . import lib, ops  # noqa: F401

__all__ = ["lib", "ops", "BertModel", "patch_transformers"]


def __getattr__(name):  # lazy: importing transformers costs seconds
    if name in ("BertModel", "patch_transformers"):
        from . import modeling_bert
        return getattr(modeling_bert, name)
    raise AttributeError(name)
