# config.py -- same names as the reference's image_processor/config.py (Config, load_config), so the drop-in stage
# scripts also run stand-alone from this directory.  When the scripts are copied into the reference tree, the
# reference's own config.py is used instead.
import _omni_path

_omni_path.add()
from omni_b200.config import Config, load_config, HOT_PATH_KEYS  # noqa: E402,F401
