"""config.json contract of the pipeline (mirror of the reference's config.py:10-132).

The key names and defaults ARE the drop-in contract: pipeline.py serialises `Config()` into
<out>/config.json (pipeline.py:21-45) and every stage reads it back through load_config() with the
CONFIG_PATH environment variable (config.py:107-132; unknown keys dropped, raw dict kept in _raw).
The table below lists (key, default, group); stages 01-03 read only the groups "io", "color", "edge".
"""
from __future__ import annotations

import copy
import dataclasses
import json
import os
from typing import Any

_FIELDS: list[tuple[str, Any, str]] = [
    ("input_image", "input.png", "io"),
    ("output_dir", "output", "io"),
    ("n_cores", 12, "io"),
    ("max_dimension", 2000, "io"),
    ("color_names", ["layer_dark", "layer_mid", "layer_skin", "layer_light"], "color"),
    ("colors", [(0, 0, 0), (255, 0, 0), (0, 255, 0), (0, 0, 255)], "color"),
    ("color_tolerance", 30, "color"),
    ("edge_low_threshold", 50, "edge"),
    ("edge_high_threshold", 150, "edge"),
    ("edge_kernel_size", 3, "edge"),
    ("edge_morph_kernel", 3, "edge"),
    ("edge_morph_open_iters", 1, "edge"),
    ("edge_morph_close_iters", 1, "edge"),
    ("smoothing_iterations", 2, "edge"),
    # --- downstream stages 04-13 (not computed here; kept so one config.json serves the whole run) ---
    ("min_contour_area", 10.0, "vector"),
    ("epsilon_factor", 0.002, "vector"),
    ("dedup_max_passes", 10, "vector"),
    ("target_width_mm", 210, "plot"),
    ("target_height_mm", 297, "plot"),
    ("pixels_per_mm", 40, "plot"),
    ("margin_left_mm", 10.0, "plot"),
    ("margin_right_mm", 10.0, "plot"),
    ("margin_top_mm", 10.0, "plot"),
    ("margin_bottom_mm", 10.0, "plot"),
    ("pen_width_px", 60, "plot"),
    ("pen_radius_px", 30, "plot"),
    ("tap_max_area", 1200.0, "tap"),
    ("tap_max_perimeter", 160.0, "tap"),
    ("tap_max_dim", 25, "tap"),
    ("tap_merge_radius_px", 30, "tap"),
    ("thinning_min_segment_len", 5, "vector"),
    ("thinning_dt_margin", 0.0, "vector"),
    ("dedup_sample_step", 8, "dedup"),
    ("dedup_overlap_threshold", 0.60, "dedup"),
    ("dedup_draw_antialiased", False, "dedup"),
    ("ignore_tail_points_intra", 120, "dedup"),
    ("collision_radius_intra_px", 18.0, "dedup"),
    ("collision_radius_global_px", 21.0, "dedup"),
    ("hash_stride_px", 18.0, "dedup"),
    ("max_join_jump_px", 80.0, "dedup"),
    ("simplify_enabled", False, "vector"),
    ("stop_after_edges", False, "io"),
    ("stream_force_color_index", None, "stream"),
    ("stream_color_by_name", None, "stream"),
    ("stream_color_by_order", None, "stream"),
]


def _default_factory(v):
    return (lambda: copy.deepcopy(v))


def _ensure_output_dirs(self) -> None:
    """config.py:93-97: <out>/ and one directory per colour layer."""
    os.makedirs(self.output_dir, exist_ok=True)
    for layer in self.color_names:
        os.makedirs(os.path.join(self.output_dir, layer), exist_ok=True)


Config = dataclasses.make_dataclass(
    "Config",
    [(k, Any, dataclasses.field(default_factory=_default_factory(v)) if isinstance(v, (list, dict))
      else dataclasses.field(default=v)) for k, v, _g in _FIELDS],
    namespace={"ensure_output_dirs": _ensure_output_dirs, "__doc__": "Pipeline configuration (see module docstring)."},
)

HOT_PATH_KEYS = [k for k, _v, g in _FIELDS if g in ("io", "color", "edge")]


def load_config(path: str | None = None):
    """Read <path> or $CONFIG_PATH; keep only known keys; defaults on any failure (config.py:107-132)."""
    src = path or os.environ.get("CONFIG_PATH")
    if not src:
        return Config()
    try:
        with open(src, "r", encoding="utf-8") as fh:
            raw = json.load(fh)
    except Exception as exc:                      # same forgiving behaviour as the reference
        print(f"[config] WARNING: failed to read JSON ({exc}); using defaults.")
        return Config()
    known = {f.name for f in dataclasses.fields(Config)}
    cfg = Config(**{k: v for k, v in raw.items() if k in known})
    cfg._raw, cfg._path = raw, src
    print(f"[config] Loading config: {src} (exists=True)")
    return cfg
