"""`feather` (feather-format) as the reference uses it: cluster_helpers.py:78, :116, :205, :213;
pixel_som_clustering.py:74, :118, :134, :185."""
from ark_analysis_b200.io_utils import read_dataframe, write_dataframe  # noqa: F401
