# front_half.py -- stages 01, 02 and 03 of the reference's pipeline in ONE process (same files, same log lines): what
# `pipeline.py IMG --start-step 1 --end-step 3` produces with the drop-in stage scripts, without two of the three interpreter
# starts / CUDA contexts and without decoding resized.png again.  CONFIG_PATH as for the stage scripts.
import _omni_path

_omni_path.add()
load_config = _omni_path.load_config_fn()
from omni_b200 import stages  # noqa: E402


def main():
    stages.front_half_main(load_config())


if __name__ == "__main__":
    main()
