"""LocalPoolPointnet parameter container (reference shapeformer/models/vqdif/enc.py:11-64): reproduces the reference's
state_dict keys; evaluation is in shapeformer_b200.encoder.PointEncoder (csrc/enc_kernels.cu)."""
import torch.nn as nn


class _ResFC(nn.Module):
    def __init__(self, size_in, size_out):
        super().__init__()
        self.fc_0, self.fc_1 = nn.Linear(size_in, min(size_in, size_out)), nn.Linear(min(size_in, size_out), size_out)
        self.shortcut = nn.Linear(size_in, size_out, bias=False)


class _CRG(nn.Module):
    def __init__(self, ci, co, k):
        super().__init__()
        self.conv = nn.Conv3d(ci, co, k, stride=k, padding=0, bias=False)
        self.groupnorm = nn.GroupNorm(8, co)


class Downsampler(nn.Module):
    def __init__(self, in_channels, downsample_steps=1):
        super().__init__()
        ch = [in_channels * 2 ** k for k in range(downsample_steps + 1)]
        blocks = []
        for i in range(downsample_steps):
            blocks += [_CRG(ch[i], ch[i + 1], 2), _CRG(ch[i + 1], ch[i + 1], 1)]
        self.blocks = nn.Sequential(*blocks)


class LocalPoolPointnet(nn.Module):
    def __init__(self, c_dim=128, dim=3, hidden_dim=128, scatter_type="max", downsampler=False, downsampler_kwargs=None,
                 c2i_order="original", grid_resolution=None, plane_type="grid", padding=0.1, n_blocks=5):
        super().__init__()
        shipped = (c_dim, dim, hidden_dim, scatter_type, c2i_order, grid_resolution, padding, n_blocks) == \
                  (32, 3, 32, "max", "original", 64, 0.1, 5) and "grid" in plane_type and downsampler and \
                  dict(downsampler_kwargs) == dict(in_channels=32, downsample_steps=2)
        if not shipped:
            raise NotImplementedError("only the shipped LocalPoolPointnet (hidden = c_dim = 32, 64^3 grid, 'max', 2 downsample steps)")
        self.fc_pos = nn.Linear(dim, 2 * hidden_dim)
        self.blocks = nn.ModuleList([_ResFC(2 * hidden_dim, hidden_dim) for _ in range(n_blocks)])
        self.fc_c = nn.Linear(hidden_dim, c_dim)
        self.downsampler = Downsampler(**downsampler_kwargs)
