"""Locates the omni_b200 package for the drop-in stage scripts.

The scripts are meant to be copied next to the reference's pipeline.py (which runs `<dir of pipeline.py>/NN_name.py`
as a subprocess, pipeline.py:55-64,88-98).  Set OMNI_B200_HOME to the directory that contains the `omni_b200/`
package (this repository's `omnirevolve-image-processor_b200/`) unless the scripts are run from inside this tree.
"""
import os
import sys


def add():
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("OMNI_B200_HOME"), os.path.dirname(here), here):
        if cand and os.path.isdir(os.path.join(cand, "omni_b200")):
            if cand not in sys.path:
                sys.path.insert(0, cand)
            return cand
    raise ImportError("omni_b200 package not found: set OMNI_B200_HOME to <repo>/omnirevolve-image-processor_b200")


def load_config_fn():
    """The reference's own config.load_config when its config.py sits beside the scripts, else the mirror."""
    try:
        from config import load_config          # reference image_processor/config.py
    except ImportError:
        from omni_b200.config import load_config
    return load_config
