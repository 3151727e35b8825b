"""CPU oracle for the Pixie SOM hot path -- TEST INFRASTRUCTURE ONLY (parity unpinned).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package.  The product package ``ark_analysis_b200`` never does.  See ``oracle/pixie_oracle.c`` for
the provenance of every function (reference file:line) and why parity is "unpinned".
"""
from .pixie_oracle import (  # noqa: F401
    build, grid_chebyshev, map_data_to_nodes, map_data_to_nodes_f32, map_data_to_nodes_mt,
    som_online, som_batch, cluster_sums_f32, default_radius, init_codebook_indices, TILE,
)
