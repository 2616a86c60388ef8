"""Drop-in for subphaser/Data.py: the `.kmer.mat` text loader (Data.py:6-21), same attributes
(`data`, `colnames`, `rownames`, `d_rows`), parsed with pandas' C reader (round-trip float parsing
gives the same doubles as Python's float())."""
import numpy as np


class LoadData:
    def __init__(self, datafile):
        self.datafile = datafile

    def load_matrix(self):
        import pandas as pd
        with open(self.datafile) as f:
            header = f.readline().strip().split()
        self.colnames = header[1:]
        ncol = len(self.colnames)
        try:
            df = pd.read_csv(self.datafile, sep=r"\s+", header=None, skiprows=1, engine="c",
                             float_precision="round_trip", dtype={0: str}, na_filter=False)
        except pd.errors.EmptyDataError:
            df = pd.DataFrame({i: [] for i in range(ncol + 1)})
        self.rownames = df[0].astype(str).tolist()
        self.data = np.ascontiguousarray(df.iloc[:, 1:1 + ncol].to_numpy(dtype=np.float64))
        if self.data.ndim != 2:
            self.data = self.data.reshape(len(self.rownames), ncol)
        self._d_rows = None

    @property
    def d_rows(self):
        if self._d_rows is None:
            self._d_rows = dict(zip(self.rownames, self.data.tolist()))
        return self._d_rows
