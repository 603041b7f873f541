"""Build an experimental variant of the library next to the product one:  python tools/ab_build.py NAME [DEFINE=VALUE ...]
-> omnirevolve-image-processor_b200/lib/libomni_NAME.so (git-ignored, travels to the GPU box)."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, "omnirevolve-image-processor_b200"))
import build  # noqa: E402

name, defs = sys.argv[1], tuple(sys.argv[2:])
print(build.build(out=os.path.join(build.LIB_DIR, f"libomni_{name}.so"), defines=defs))
