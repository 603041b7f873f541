#!/usr/bin/env python3
# process_colors.py -- drop-in for the reference's stand-alone process_colors.py (v1.1.1): same command line, same files in the
# output directory (labels.png, labels.npy, palette.json, layer_<i>_<name>.png), same log lines.  The palette still comes from
# cv2.kmeans or a palette JSON on the host; assign_labels (process_colors.py:69-77, int16 wrap reproduced), the class histogram and
# the one-hot layers run on the GPU (libomni_b200).  No CPU fallback.
import _omni_path

_omni_path.add()
from omni_b200 import colors_cli  # noqa: E402

load_image_rgb = colors_cli.load_image_rgb
kmeans_palette = colors_cli.kmeans_palette
palette_from_json = colors_cli.palette_from_json
assign_labels = colors_cli.assign_labels
default_color_names = colors_cli.default_color_names
save_labels_png = colors_cli.save_labels_png
main = colors_cli.main

if __name__ == "__main__":
    main()
