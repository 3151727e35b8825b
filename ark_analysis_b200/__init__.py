"""pixie-b200: the Pixie SOM hot path of angelolab/ark-analysis on NVIDIA B200 (sm_100a).

Public surface (mirrors the reference for this path and nothing else):

* ``ark_analysis_b200.som``                  -- device operators + the pyFlowSOM-shaped ``som`` /
  ``map_data_to_nodes`` (reference: cluster_helpers.py:14, :106-109, :152-157)
* ``ark_analysis_b200.cluster_helpers``      -- ``PixieSOMCluster`` / ``PixelSOMCluster`` /
  ``CellSOMCluster`` (reference: cluster_helpers.py:52-416)
* ``ark_analysis_b200.pixel_som_clustering`` -- ``train_pixel_som`` / ``cluster_pixels`` /
  ``generate_som_avg_files`` (reference: pixel_som_clustering.py)
* ``ark_analysis_b200.cell_som_clustering``  -- ``train_cell_som`` / ``cluster_cells`` /
  ``generate_som_avg_files`` (reference: cell_som_clustering.py)
* ``ark_analysis_b200.compat.install()``     -- registers ``ark.phenotyping.*`` / ``pyFlowSOM``
  aliases so notebook cells written against the reference run unchanged.
"""
__version__ = "0.1.0"
