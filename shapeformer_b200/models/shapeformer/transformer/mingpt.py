"""CondTupleGPT with the reference's constructor, state_dict keys and method surface (reference
shapeformer/models/shapeformer/transformer/mingpt.py:185-319), evaluated by libsfb200's KV-cached CUDA engine.

The nn.Module tree below is a parameter CONTAINER: it reproduces the reference's parameter names/shapes
so reference checkpoints load with load_state_dict; no torch op of it is ever executed.  Class name contains "TupleGPT"
(asserted by ShapeFormer.__init__, shapeformer.py:24).
"""
import torch
import torch.nn as nn

from .... import ar as _ar


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.key, self.query, self.value, self.proj = (nn.Linear(d, d) for _ in range(4))


class _Block(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.ln1, self.ln2 = nn.LayerNorm(d), nn.LayerNorm(d)
        self.attn = _Attn(d)
        self.mlp = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Linear(4 * d, d), nn.Dropout(0.0))


class CondTupleGPT(nn.Module):
    def __init__(self, vocab_sizes, extra_vocab_sizes, block_size, tuple_n, n_layers=(12,), n_head=8, n_embd=256,
                 embd_pdrop=0., resid_pdrop=0., attn_pdrop=0., n_unmasked=0, no_pos_emb=False, cond_emb_same=False,
                 pos_no_restart=False, head_hidden_layers=0):
        super().__init__()
        if tuple_n != 2 or len(vocab_sizes) != 2 or len(n_layers) != 2 or len(extra_vocab_sizes) != 1:
            raise NotImplementedError("B200 path covers the shipped (pos, val) tuple model with one extra index")
        if n_unmasked or no_pos_emb or cond_emb_same or pos_no_restart or head_hidden_layers:
            raise NotImplementedError("only the shipped configuration of CondTupleGPT is on the B200 path")
        if n_embd != 64 * n_head:
            raise NotImplementedError("head dim must be 64")
        self.spec = dict(n_embd=n_embd, n_head=n_head, n_layers=tuple(n_layers), block_size=block_size,
                         vocab_sizes=tuple(vocab_sizes), extra_vocab_sizes=tuple(extra_vocab_sizes))
        self.tuple_n, self.block_size = tuple_n, block_size
        self.tok_embs = nn.ModuleList([nn.Embedding(v, n_embd) for v in vocab_sizes])
        self.extra_tok_embs = nn.ModuleList([nn.Embedding(v, n_embd) for v in extra_vocab_sizes])
        self.blocks = nn.ModuleList([nn.Sequential(*[_Block(n_embd) for _ in range(n)]) for n in n_layers])
        self.heads = nn.ModuleList([nn.Sequential(nn.LayerNorm(n_embd), nn.Linear(n_embd, v, bias=False))
                                    for v in vocab_sizes])
        self.pos_emb = nn.Parameter(torch.zeros(1, block_size, n_embd))
        self.cond_pos_emb = nn.Parameter(torch.zeros(1, block_size, n_embd))
        self.apply(self._init_weights)
        self.requires_grad_(False)
        self.eval()
        self._packed = None
        self._samplers = {}
        # reference checkpoints also carry the causal-mask buffers (blocks.g.l.attn.mask): accept and drop them
        self._register_load_state_dict_pre_hook(self._drop_masks)
        # fires for direct AND nested loads (a parent ShapeFormer.load_state_dict recurses through
        # _load_from_state_dict, never through this module's load_state_dict): new weights invalidate the packed blob,
        # the samplers built on it and their captured graphs
        self.register_load_state_dict_post_hook(self._invalidate)

    @staticmethod
    def _invalidate(module, incompatible_keys):
        module._packed, module._samplers = None, {}

    @staticmethod
    def _drop_masks(state_dict, prefix, *args):
        for k in [k for k in state_dict if k.startswith(prefix) and k.endswith("attn.mask")]:
            del state_dict[k]

    @staticmethod
    def _init_weights(m):   # reference init (mingpt.py:248-255)
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)

    def get_block_size(self):
        return self.block_size

    # ------------------------------------------------------------------------------------------------------------------
    def packed_weights(self):
        """fp32 weight blob in the library's layout on this module's CUDA device (built once, after loading)."""
        dev = self.pos_emb.device
        if dev.type != "cuda":
            raise RuntimeError("CondTupleGPT (B200) runs on a CUDA device only — move the model with .cuda(); "
                               "there is no CPU fallback")
        if self._packed is None or self._packed.device != dev:
            self._packed = _ar.pack_gpt_weights(self.state_dict(), self.spec, dev)
            self._samplers = {}
        return self._packed

    def sampler(self, rows, L_cond, max_steps, end_tokens, keep_history=True):
        """An ARSampler with CAPACITY for this batch (cached and reused): the conditioning length is bucketed to a multiple
        of 64 so that shapes with different L_cond share one KV cache, one workspace and one captured step graph (the
        sampler accepts any L_c <= max_cond, B <= max_rows)."""
        cap = min(-(-max(int(L_cond), 1) // 64) * 64, self.block_size - 1)
        key = (rows, cap, max_steps, tuple(end_tokens), keep_history)
        s = self._samplers.get(key)
        if s is None:
            self._samplers.clear()   # one live KV cache at a time
            s = _ar.ARSampler(self.packed_weights(), self.spec, end_tokens, max_rows=rows, max_cond=cap,
                              max_steps=max_steps, keep_history=keep_history)
            self._samplers[key] = s
        return s

    def sample_next_tuple(self, idx, extra_idx=None, L_cond=1):
        raise NotImplementedError("the uncached generator protocol (mingpt.py:297-310) is replaced by the KV-cached engine: "
                                  "use ShapeFormer.sample / sample_indices")

    def forward(self, idx, extra_idx=None, L_cond=1, target_idx=None):
        raise NotImplementedError("teacher-forced forward (training) is outside the B200 hot path")
