from ark_analysis_b200.io_utils import (list_files, remove_file_extensions,  # noqa: F401
                                        validate_paths)


def list_folders(dir_name, substrs=None, exact_match=False, ignore_hidden=True):
    import os
    names = [f for f in os.listdir(dir_name) if os.path.isdir(os.path.join(dir_name, f))]
    if ignore_hidden:
        names = [f for f in names if not f.startswith(".")]
    if substrs is not None:
        if isinstance(substrs, str):
            substrs = [substrs]
        names = [f for f in names
                 if any((f == s) if exact_match else (s in f) for s in substrs)]
    return sorted(names)
