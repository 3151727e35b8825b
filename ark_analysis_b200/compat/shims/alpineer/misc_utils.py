from ark_analysis_b200.io_utils import verify_in_list, verify_same_elements  # noqa: F401


def make_iterable(a, ignore_str=True):
    if isinstance(a, str) and ignore_str:
        return [a]
    return a if hasattr(a, "__iter__") else [a]
