# 02_color_extract.py -- drop-in for the reference's stage 02 (k-means Lab mode, the only mode reachable through
# config.json).  cv2.kmeans stays on the host exactly as in the reference; the per-pixel Lab conversion,
# nearest-centre assignment, one-hot masks and RECT-3 open/close run on the GPU (libomni_b200).  No CPU fallback.
import _omni_path

_omni_path.add()
load_config = _omni_path.load_config_fn()
from omni_b200 import stages  # noqa: E402

_darkness_rank = stages._darkness_rank
_ensure_bgr = stages._ensure_bgr
_kmeans_lab = stages._kmeans_lab
_lab_to_bgr = stages._lab_to_bgr


def main():
    stages.color_extract_main(load_config())


if __name__ == "__main__":
    main()
