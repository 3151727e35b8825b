"""TEST-ONLY stand-in for `pyFlowSOM` backed by the CPU oracle (oracle/pixie_oracle.c): lets the
reference's own hot-path tests run, unchanged, against the ORACLE on a box without a GPU -- which
pins the oracle on every invariant those tests assert.  Never on the product path."""
import numpy as np

import oracle


def som(data, xdim=10, ydim=10, rlen=10, alpha_range=(0.05, 0.01), radius_range=None, distf=2,
        nodes=None, importance=None, seed=None):
    return oracle.som_online(np.ascontiguousarray(data, np.float64), xdim, ydim, rlen=rlen,
                             alpha_range=alpha_range, seed=seed)


def map_data_to_nodes(nodes, newdata, distf=2):
    return oracle.map_data_to_nodes(np.ascontiguousarray(nodes, np.float64),
                                    np.ascontiguousarray(newdata, np.float64))
