"""Quantizer container (reference shapeformer/models/vqdif/quantizer.py): only get_code is on the hot path."""
import torch
import torch.nn as nn


class Quantizer(nn.Module):
    def __init__(self, vocab_size, n_embd, gamma=0.99, x_dim=3):
        super().__init__()
        self.embedding = nn.Embedding(vocab_size, n_embd)
        self.embedding.weight.requires_grad = False
        self.n_embd, self.vocab_size, self.gamma, self.x_dim = n_embd, vocab_size, gamma, x_dim
        self.register_buffer("N", torch.zeros(vocab_size))
        self.register_buffer("z_avg", self.embedding.weight.data.clone())

    def forward(self, grid_feat):
        raise NotImplementedError("nearest-code search belongs to the encoder side ('next' row §8f-1)")
