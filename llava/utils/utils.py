def disable_torch_init():
    """llava/utils/utils.py:110-117 skips torch's default nn.Linear / LayerNorm initialisation to speed up model
    construction.  Our model never builds nn.Modules (weights are loaded straight into kernel layouts), so this is a
    no-op kept for call-site compatibility (eval_spatial.py:111)."""
    return None
