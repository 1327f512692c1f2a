"""LocalDecoder parameter container (reference shapeformer/models/vqdif/dec.py, unet3d.py, updown.py): reproduces the
reference's state_dict keys; evaluation is in shapeformer_b200.decoder.ImplicitDecoder."""
import torch.nn as nn


class _GCR(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.groupnorm = nn.GroupNorm(8, ci)
        self.conv = nn.Conv3d(ci, co, 3, padding=1, bias=False)


class _CRG(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.conv = nn.Conv3d(ci, co, 3, padding=1, bias=False)
        self.groupnorm = nn.GroupNorm(8, co)


class _Double(nn.Module):
    def __init__(self, ci, cm, co):
        super().__init__()
        self.SingleConv1, self.SingleConv2 = _GCR(ci, cm), _GCR(cm, co)


class _Level(nn.Module):
    def __init__(self, ci, cm, co):
        super().__init__()
        self.basic_module = _Double(ci, cm, co)


class UNet3D(nn.Module):
    def __init__(self, in_channels, out_channels, f_maps=64, num_levels=4, **kwargs):
        super().__init__()
        if num_levels != 3 or in_channels != f_maps or out_channels != f_maps or kwargs.get("layer_order", "gcr") != "gcr":
            raise NotImplementedError("only the shipped UNet3D (3 levels, 'gcr', in = out = f_maps) is supported")
        f = f_maps
        self.encoders = nn.ModuleList([_Level(f, f, f), _Level(f, f, 2 * f), _Level(2 * f, 2 * f, 4 * f)])
        self.decoders = nn.ModuleList([_Level(6 * f, 2 * f, 2 * f), _Level(3 * f, f, f)])
        self.final_conv = nn.Conv3d(f, out_channels, 1)


class Upsampler(nn.Module):
    def __init__(self, in_channels, upsampler_steps=1, mode="nearest"):
        super().__init__()
        if upsampler_steps != 2 or mode != "nearest":
            raise NotImplementedError("only the shipped Upsampler (2 nearest steps) is supported")
        c = [in_channels, in_channels // 2, in_channels // 4]
        blocks = []
        for s in range(2):
            blocks += [nn.Identity(), _CRG(c[s], c[s + 1]), _CRG(c[s + 1], c[s + 1])]
        self.blocks = nn.Sequential(*blocks)


class _ResFC(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.fc_0, self.fc_1 = nn.Linear(h, h), nn.Linear(h, h)


class LocalDecoder(nn.Module):
    def __init__(self, dim=3, c_dim=128, unet3d=False, unet3d_kwargs=None, upsampler=False, upsampler_kwargs=None,
                 hidden_size=256, n_blocks=5, leaky=False, sample_mode="bilinear", padding=0.1):
        super().__init__()
        if (dim, c_dim, hidden_size, n_blocks, leaky, sample_mode, padding) != (3, 32, 32, 5, False, "bilinear", 0.1) \
                or not unet3d or not upsampler:
            raise NotImplementedError("only the shipped LocalDecoder (c_dim = hidden = 32, 5 blocks, UNet3D + Upsampler)")
        self.unet3d = UNet3D(**unet3d_kwargs)
        self.upsampler = Upsampler(**upsampler_kwargs)
        self.fc_c = nn.ModuleList([nn.Linear(c_dim, hidden_size) for _ in range(n_blocks)])
        self.fc_p = nn.Linear(dim, hidden_size)
        self.blocks = nn.ModuleList([_ResFC(hidden_size) for _ in range(n_blocks)])
        self.fc_out = nn.Linear(hidden_size, 1)
