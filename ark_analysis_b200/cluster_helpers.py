"""SOM cluster objects of the Pixie path -- the host-side mirror of
``/root/reference/src/ark/phenotyping/cluster_helpers.py:52-416`` (``PixieSOMCluster``,
``PixelSOMCluster``, ``CellSOMCluster``): same constructor arguments and defaults, same attributes
(``weights``, ``norm_data``, ``train_data``, ``cell_data``, ``som_clusters_seen``), same warning
strings and errors, same files written.  What differs is underneath: the two pyFlowSOM calls
(reference lines 106-109 and 152-157) are replaced by the B200 operators of ``som.py`` -- training
flattens the table into one contiguous fp32 device matrix and runs the batch SOM on it, assignment
streams the rows through the tensor-core BMU kernel.  There is no CPU path.
"""
import os
import pathlib
import warnings
from abc import ABC, abstractmethod
from typing import List

import numpy as np
import pandas as pd

from . import io_utils, som
from .io_utils import list_files, validate_paths, verify_in_list

__all__ = ["PixieSOMCluster", "PixelSOMCluster", "CellSOMCluster"]


class PixieSOMCluster(ABC):
    """Generic SOM runner (reference: cluster_helpers.py:52-163)."""

    @abstractmethod
    def __init__(self, weights_path: pathlib.Path, columns: List[str], num_passes: int = 1,
                 xdim: int = 10, ydim: int = 10, lr_start: float = 0.05, lr_end: float = 0.01,
                 seed=42):
        self.weights_path = weights_path
        # the weights file doubles as the training checkpoint (reference :77-80)
        self.weights = io_utils.read_dataframe(weights_path) if os.path.exists(weights_path) \
            else None
        self.columns = columns
        self.num_passes = num_passes
        self.xdim = xdim
        self.ydim = ydim
        self.lr_start = lr_start
        self.lr_end = lr_end
        self.seed = seed
        # device the kernels run on (None = current CUDA device); not part of the reference API
        self.device = None

    @abstractmethod
    def normalize_data(self) -> pd.DataFrame:
        """Normalisation applied to the input table before it reaches the SOM."""

    def train_som(self, data: pd.DataFrame):
        """Train the SOM on ``data`` and save the weights (reference :98-116).

        ``data`` [n, C] is flattened to the fp32 device matrix; the codebook comes back as float64
        [xdim*ydim, C] exactly like ``pyFlowSOM.som`` returns it."""
        som_weights = som.som(
            data=data.values, xdim=self.xdim, ydim=self.ydim, rlen=self.num_passes,
            alpha_range=(self.lr_start, self.lr_end), seed=self.seed, device=self.device)
        som_weights = np.reshape(som_weights, (self.xdim * self.ydim, som_weights.shape[-1]))
        self.weights = pd.DataFrame(som_weights, columns=list(data.columns))
        io_utils.write_dataframe(self.weights, self.weights_path, compression='uncompressed')

    def generate_som_clusters(self, external_data: pd.DataFrame,
                              num_parallel_obs: int = 1000000) -> np.ndarray:
        """SOM cluster label (1-indexed) of every row of ``external_data`` (reference :118-163).

        ``num_parallel_obs`` keeps its meaning as the number of rows handed to the kernel at a
        time (it sizes the H2D / kernel / D2H pipeline chunks); labels do not depend on it."""
        if num_parallel_obs <= 0:
            raise ValueError("num_parallel_obs specified needs to be greater than 0")

        weights_cols = list(self.weights.columns)
        verify_in_list(weights_cols=weights_cols, external_data_cols=external_data.columns.values)

        # no rows (an image without pixels): same sentinel as the reference
        if external_data.shape[0] == 0:
            return np.empty(0)

        # column order follows the weights, like the reference's `.loc[..., weights_cols]`
        values = external_data[weights_cols].to_numpy()
        if values.dtype not in (np.float32, np.float64):
            values = values.astype(np.float64)
        labels, _ = som.map_data_to_nodes(
            self.weights.values.astype(values.dtype, copy=False), values, device=self.device,
            chunk_rows=int(num_parallel_obs), return_dists=False)
        return labels


class PixelSOMCluster(PixieSOMCluster):
    """Pixel SOM (reference: cluster_helpers.py:166-301)."""

    def __init__(self, pixel_subset_folder: pathlib.Path, norm_vals_path: pathlib.Path,
                 weights_path: pathlib.Path, fovs: List[str], columns: List[str],
                 num_passes: int = 1, xdim: int = 10, ydim: int = 10,
                 lr_start: float = 0.05, lr_end: float = 0.01, seed=42):
        super().__init__(weights_path, columns, num_passes, xdim, ydim, lr_start, lr_end, seed)

        validate_paths([norm_vals_path, pixel_subset_folder])
        self.norm_data = io_utils.read_dataframe(norm_vals_path)
        self.fovs = fovs

        # training table = the subsetted pixels of the requested FOVs, in (sorted) file order
        wanted = set(fovs)
        fov_files = [f for f in list_files(pixel_subset_folder, substrs='.feather')
                     if os.path.splitext(f)[0] in wanted]
        self.train_data = pd.concat(
            [io_utils.read_dataframe(os.path.join(pixel_subset_folder, f)) for f in fov_files])
        self.train_data = self.normalize_data(self.train_data)

        self.som_clusters_seen = set()

    def normalize_data(self, external_data: pd.DataFrame) -> pd.DataFrame:
        """Divide the channel columns by the 1 x C normalisation row (reference :223-248)."""
        norm_cols = list(self.norm_data.columns)  # a list: pandas 3 string arrays do not index
        verify_in_list(norm_data_cols=norm_cols, external_data_cols=external_data.columns.values)
        normalized = external_data.copy()
        normalized[norm_cols] = normalized[norm_cols].div(self.norm_data.iloc[0], axis=1)
        return normalized

    def train_som(self, overwrite=False):
        """Train on ``train_data`` unless valid weights already exist (reference :250-268)."""
        if overwrite:
            warnings.warn('Overwrite flag set, retraining SOM')
        elif self.weights is not None:
            if set(self.weights.columns.values) == set(self.columns):
                warnings.warn('Pixel SOM already trained on specified markers')
                return
            warnings.warn('New markers specified, retraining')
        super().train_som(self.train_data[self.columns])

    def assign_som_clusters(self, external_data: pd.DataFrame, normalize_data: bool = True,
                            num_parallel_pixels: int = 1000000) -> pd.DataFrame:
        """Label every pixel of ``external_data`` (reference :270-301).  Returns the (normalised)
        table with a ``pixel_som_cluster`` column and records the clusters seen."""
        labelled = self.normalize_data(external_data) if normalize_data else external_data.copy()
        som_labels = super().generate_som_clusters(labelled, num_parallel_obs=num_parallel_pixels)
        labelled['pixel_som_cluster'] = som_labels
        self.som_clusters_seen.update(np.unique(som_labels).tolist())
        return labelled


    def assign_som_clusters_table(self, table, normalize_data: bool = True):
        """Arrow-native twin of :meth:`assign_som_clusters` for a FOV read with
        ``io_utils.read_table``.

        The float64 channel columns go to the device as the Arrow buffers they are; one kernel
        (``som.columns_to_rows``) normalises, casts and transposes them into the fp32 matrix the
        BMU kernel streams -- no DataFrame copy, no ``.loc`` gather, no float64 staging matrix.
        Returns a ``pyarrow.Table`` with the normalised channel columns and the int32
        ``pixel_som_cluster`` column, i.e. exactly the table the DataFrame path writes; ``None``
        when the table does not qualify (no rows, nulls, channels that are not float64), in which
        case the caller takes the DataFrame path."""
        import pyarrow as pa
        import torch

        names = table.column_names
        norm_cols = list(self.norm_data.columns)
        weights_cols = list(self.weights.columns)
        if normalize_data:
            verify_in_list(norm_data_cols=norm_cols, external_data_cols=names)
        verify_in_list(weights_cols=weights_cols, external_data_cols=names)
        n = table.num_rows
        if n == 0 or 'pixel_som_cluster' in names:
            return None
        touched = list(dict.fromkeys(weights_cols + (norm_cols if normalize_data else [])))
        host = {}
        for col in touched:
            arr = table.column(col)
            if arr.type != pa.float64() or arr.null_count:
                return None
            host[col] = arr.combine_chunks().to_numpy(zero_copy_only=True) \
                if arr.num_chunks != 1 else arr.chunk(0).to_numpy(zero_copy_only=True)

        norm_row = self.norm_data.iloc[0]
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None \
            else torch.device(self.device)
        C = len(weights_cols)
        cols = torch.empty((C, n), dtype=torch.float64, device=dev)
        div = np.ones(C, dtype=np.float64)
        for j, col in enumerate(weights_cols):
            with warnings.catch_warnings():  # Arrow buffers are read-only; they are only read here
                warnings.simplefilter("ignore", UserWarning)
                src = torch.from_numpy(host[col])
            cols[j].copy_(src, non_blocking=True)
            if normalize_data and col in norm_row.index:
                div[j] = float(norm_row[col])
        X = som.columns_to_rows(cols, torch.from_numpy(div) if normalize_data else None)
        W = torch.from_numpy(self.weights.values.astype(np.float64)).to(dev).float().contiguous()
        labels = som.bmu(X, W).cpu().numpy()

        out = table
        if normalize_data:
            for col in norm_cols:
                i = out.column_names.index(col)
                out = out.set_column(i, out.schema.field(i),
                                     pa.array(np.divide(host[col], float(norm_row[col]))))
        out = out.append_column('pixel_som_cluster', pa.array(labels, type=pa.int32()))
        self.som_clusters_seen.update(np.unique(labels).tolist())
        return out


class CellSOMCluster(PixieSOMCluster):
    """Cell SOM (reference: cluster_helpers.py:304-416)."""

    def __init__(self, cell_data: pd.DataFrame, weights_path: pathlib.Path, fovs: List[str],
                 columns: List[str], num_passes: int = 1, xdim: int = 10, ydim: int = 10,
                 lr_start: float = 0.05, lr_end: float = 0.01, seed=42, normalize=True):
        super().__init__(weights_path, columns, num_passes, xdim, ydim, lr_start, lr_end, seed)
        self.fovs = fovs
        # keep only the requested FOVs; positional index from here on (reference :347-349)
        self.cell_data = cell_data[cell_data['fov'].isin(self.fovs)].reset_index(drop=True)
        if normalize:
            self.normalize_data()

    def normalize_data(self):
        """99.9th-percentile normalisation of the training columns, zeros ignored
        (reference :355-372)."""
        sub = self.cell_data[self.columns].copy()
        norm_vals = sub.replace(0, np.nan).quantile(q=0.999, axis=0)
        self.cell_data[self.columns] = sub.div(norm_vals)

    def train_som(self, overwrite=False):
        """Train on ``cell_data`` unless valid weights already exist (reference :374-393)."""
        if overwrite:
            warnings.warn('Overwrite flag set, retraining SOM')
        elif self.weights is not None:
            if set(self.weights.columns.values) == set(self.columns):
                warnings.warn('Cell SOM already trained on specified columns')
                return
            warnings.warn('New columns specified, retraining')
        super().train_som(self.cell_data[self.columns])

    def assign_som_clusters(self, num_parallel_cells=1000000) -> pd.DataFrame:
        """Label every cell (reference :395-416); the table is already normalised."""
        som_labels = super().generate_som_clusters(
            self.cell_data[self.columns], num_parallel_obs=num_parallel_cells)
        self.cell_data['cell_som_cluster'] = som_labels
        return self.cell_data
