"""Drop-in aliases: after ``install()`` the module names the reference's notebooks and scripts
import resolve to this package --

* ``pyFlowSOM`` (``som``, ``map_data_to_nodes``)  -> ``ark_analysis_b200.som``
* ``ark.phenotyping.cluster_helpers`` / ``pixel_som_clustering`` / ``cell_som_clustering`` /
  ``pixel_cluster_utils`` / ``cell_cluster_utils`` / ``pixie_preprocessing`` and
  ``ark.utils.data_utils`` -> the modules of the same name here (each holds only the functions of
  the Pixie SOM path: SURVEY.md section 8)
* ``feather`` (``read_dataframe``, ``write_dataframe``) -> ``ark_analysis_b200.io_utils``

so ``templates/2_Pixie_Cluster_Pixels.ipynb`` cells 32/35 and the cell-clustering notebook run
unchanged.  Nothing is aliased unless ``install()`` is called, and an already-importable real
module is never shadowed unless ``force=True``.
"""
import importlib
import importlib.util
import sys
import types

_ARK_MODULES = ["cluster_helpers", "pixel_som_clustering", "cell_som_clustering",
                "pixel_cluster_utils", "cell_cluster_utils", "pixie_preprocessing"]
_ARK_UTILS = ["data_utils"]


def _has_real(name):
    try:
        return importlib.util.find_spec(name) is not None
    except (ImportError, ValueError):
        return False


def shim_path():
    """Directory of the stand-in modules (feather, natsort, alpineer, skimage.io, pyFlowSOM) that
    let the UNMODIFIED reference modules import in an image without those packages.  Append it to
    ``sys.path`` (after site-packages, so a really installed package wins)."""
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def install(force=False):
    """Register the aliases in ``sys.modules``.  Returns the list of names registered."""
    from .. import io_utils, som
    done = []

    if force or not _has_real("pyFlowSOM"):
        m = types.ModuleType("pyFlowSOM")
        m.som = som.som
        m.map_data_to_nodes = som.map_data_to_nodes
        m.__doc__ = "pyFlowSOM-shaped API served by ark_analysis_b200 (B200 kernels)"
        sys.modules["pyFlowSOM"] = m
        done.append("pyFlowSOM")

    if force or not _has_real("feather"):
        f = types.ModuleType("feather")
        f.read_dataframe = io_utils.read_dataframe
        f.write_dataframe = io_utils.write_dataframe
        sys.modules["feather"] = f
        done.append("feather")

    if force or not _has_real("ark"):
        ark = sys.modules.get("ark") or types.ModuleType("ark")
        ark.__path__ = []
        phen = sys.modules.get("ark.phenotyping") or types.ModuleType("ark.phenotyping")
        phen.__path__ = []
        ark.phenotyping = phen
        sys.modules["ark"] = ark
        sys.modules["ark.phenotyping"] = phen
        for name in _ARK_MODULES:
            mod = importlib.import_module(f"ark_analysis_b200.{name}")
            sys.modules[f"ark.phenotyping.{name}"] = mod
            setattr(phen, name, mod)
            done.append(f"ark.phenotyping.{name}")
        utils = sys.modules.get("ark.utils") or types.ModuleType("ark.utils")
        utils.__path__ = []
        ark.utils = utils
        sys.modules["ark.utils"] = utils
        for name in _ARK_UTILS:
            mod = importlib.import_module(f"ark_analysis_b200.{name}")
            sys.modules[f"ark.utils.{name}"] = mod
            setattr(utils, name, mod)
            done.append(f"ark.utils.{name}")
    return done
