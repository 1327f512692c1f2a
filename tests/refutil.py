"""Helpers that build the reference's own modules with synthetic weights (oracle/ref_models.py; usable where the reference
tree is importable: /root/reference in the build container, oracle/_ref on the GPU box)."""
from oracle.ref_models import (have_reference, ref_decode_index, ref_gpt, ref_shapeformer,  # noqa: F401
                               ref_vqdif_decoder, ref_vqdif_encoder)
from oracle import ref_shim  # noqa: F401,E402
