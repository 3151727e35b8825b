"""`pyFlowSOM` as the reference binds it (cluster_helpers.py:14): both callables are the B200
operators of ark_analysis_b200.som (no CPU fallback)."""
from ark_analysis_b200.som import map_data_to_nodes, som  # noqa: F401
