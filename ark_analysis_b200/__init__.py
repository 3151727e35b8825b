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
* ``ark_analysis_b200.pixie_preprocessing``   -- ``create_fov_pixel_data`` / ``preprocess_fov_device`` /
  ``fov_channel_quantiles`` (reference: pixie_preprocessing.py:18-80, :405-410)
* ``ark_analysis_b200.cell_cluster_utils``    -- ``create_c2pc_data`` (reference:
  cell_cluster_utils.py:63-192); ``ark_analysis_b200.data_utils`` -- ``generate_pixel_cluster_mask``
  (reference: utils/data_utils.py:476-555)
* ``ark_analysis_b200.compat.install()``     -- registers ``ark.phenotyping.*`` / ``pyFlowSOM``
  aliases so notebook cells written against the reference run unchanged.
"""
__version__ = "0.1.0"
