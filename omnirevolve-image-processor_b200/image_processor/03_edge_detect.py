# 03_edge_detect.py -- drop-in for the reference's stage 03: per layer ELLIPSE open/close -> GaussianBlur -> Canny,
# then the composite preview.  All layers go through one batched GPU call (libomni_b200, omni_host_edges) instead
# of the reference's process pool.  No CPU fallback.
import _omni_path

_omni_path.add()
load_config = _omni_path.load_config_fn()
from omni_b200 import stages  # noqa: E402

_ensure_odd = stages._ensure_odd
process_color = stages.process_color
detect_all_edges = stages.detect_all_edges
save_edges_composite = stages.save_edges_composite

if __name__ == "__main__":
    # detect_all_edges + save_edges_composite (03:112-115); when this tree's stage 02 has just run on the same masks and edge
    # keys, its parked planes are published instead of recomputed (omni_b200/stages.py, "hand-off 02 -> 03")
    stages.edge_detect_main(load_config())
