"""Make the reference's scripts import this implementation unchanged.

    import m3dssd_b200.dropin; m3dssd_b200.dropin.install()

registers the mirrored modules under the reference's import names -- `model.M3d_inference_align`,
`model.pose_dla_dcn`, `model.DCNv2.dcn_v2[_func]`, `model.module.attention`,
`model.module.feturealign_mgpu`, `lib.nms.gpu_nms` -- so that
`import_module('model.' + conf.model).build(conf, phase)` (scripts/test_rpn_3d.py:48, lib/core.py:69-70)
and `from lib.nms.gpu_nms import gpu_nms` (lib/rpn_util.py:16) resolve to the B200 path.  The
reference's own `lib.rpn_util`, loss, data loading and evaluation keep working on top of it.
"""
import importlib
import sys
import types

_MAP = {
    "model.M3d_inference_align": "m3dssd_b200.model.M3d_inference_align",
    "model.pose_dla_dcn": "m3dssd_b200.model.pose_dla_dcn",
    "model.DCNv2.dcn_v2": "m3dssd_b200.model.DCNv2.dcn_v2",
    "model.DCNv2.dcn_v2_func": "m3dssd_b200.model.DCNv2.dcn_v2_func",
    "model.module.attention": "m3dssd_b200.model.module.attention",
    "model.module.feturealign_mgpu": "m3dssd_b200.model.module.feturealign_mgpu",
    "lib.nms.gpu_nms": "m3dssd_b200.lib.nms.gpu_nms",
}


# opt-in: `from lib.loss.rpn_3d import *` (scripts/train_rpn_3d.py:24) then yields the static-shape RPN_3D_loss_smp.  Not in
# the default map because the reference's module also re-exports lib.rpn_util through its own star import.
_LOSS_MAP = {"lib.loss.rpn_3d": "m3dssd_b200.lib.loss.rpn_3d"}


def install(override_existing=True, loss=False):
    """Alias our modules under the reference's names.  Parent packages that are not importable
    (running outside the reference checkout) are created as empty namespace modules.  loss=True also aliases
    lib.loss.rpn_3d (the device-side RPN_3D_loss_smp)."""
    for ref_name, ours in list(_MAP.items()) + (list(_LOSS_MAP.items()) if loss else []):
        if not override_existing and ref_name in sys.modules:
            continue
        parts = ref_name.split(".")
        for i in range(1, len(parts)):
            parent = ".".join(parts[:i])
            if parent not in sys.modules:
                try:
                    importlib.import_module(parent)
                except Exception:  # noqa: BLE001
                    m = types.ModuleType(parent)
                    m.__path__ = []
                    sys.modules[parent] = m
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    return sorted(_MAP)
