# 01_resize.py -- drop-in for the reference's stage 01 (same file name, entry point, outputs and log lines);
# the INTER_AREA shrink runs on the GPU (libomni_b200, omni_resize_area_u8c3).  No CPU fallback.
import _omni_path

_omni_path.add()
load_config = _omni_path.load_config_fn()
from omni_b200 import stages  # noqa: E402


def resize_if_needed(image_path, cfg):
    return stages.resize_if_needed(image_path, cfg)


if __name__ == "__main__":
    stages.resize_main(load_config())
