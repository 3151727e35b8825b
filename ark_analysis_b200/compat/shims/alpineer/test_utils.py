from ark_analysis_b200.io_utils import make_blank_file as _make_blank_file  # noqa: F401
