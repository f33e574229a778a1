"""Result sink of the batched predict loop: one Parquet file per slab, written by a background thread from pinned,
double-buffered host copies, resumable.

The reference collects every prediction of a run in Python lists and builds one DataFrame at the end
(processors/vcfprocessor.py:261-277: `trainer.predict` -> `format_output`), which for a cohort run (2,330 genomes x
17,859 genes x 63 tissues x 1536-d embeddings, ~13 TB of fp32) is neither resumable nor able to fit in host memory.
Here a slab's results leave the GPU as soon as the slab is done:

  device results --async D2H on a copy stream--> pinned host buffer (two in rotation)
      --background thread--> slab_<index>.parquet (written to a temporary name, then renamed: a file that exists is whole)

Columns (one row per (gene, tissue) prediction, the long form of the reference's output frame):
  slab int32 | item int32 (row of the query within the slab) | sample string | gene_id string | tissue int16 |
  predicted_expression float32 | embeddings fixed_size_list<float32>[D] (omitted with embeddings=False)
`ResultWriter.done()` lists the slabs already on disk, so an interrupted run continues where it stopped.
"""
import json
import os
import queue
import threading

import numpy as np
import torch


class ResultWriter:
    def __init__(self, out_dir, emb_dim, embeddings=True, meta=None):
        import pyarrow as pa
        self.pa, self.out_dir, self.emb_dim, self.embeddings = pa, out_dir, int(emb_dim), embeddings
        os.makedirs(out_dir, exist_ok=True)
        fields = [("slab", pa.int32()), ("item", pa.int32()), ("sample", pa.string()), ("gene_id", pa.string()),
                  ("tissue", pa.int16()), ("predicted_expression", pa.float32())]
        if embeddings:
            fields.append(("embeddings", pa.list_(pa.float32(), self.emb_dim)))
        self.schema = pa.schema(fields)
        man = os.path.join(out_dir, "manifest.json")
        want = {"emb_dim": self.emb_dim, "embeddings": bool(embeddings), "meta": meta or {}}
        if os.path.exists(man):
            have = json.load(open(man))
            if have != want:
                raise ValueError(f"{out_dir} holds results of a different run ({have} != {want}); refusing to mix them")
        else:
            json.dump(want, open(man, "w"))
        self._q = queue.Queue(maxsize=2)
        self._err = None
        self._thread = threading.Thread(target=self._drain, daemon=True)
        self._thread.start()

    # -- resume ------------------------------------------------------------------------------------------------------
    def path(self, slab):
        return os.path.join(self.out_dir, f"slab_{int(slab):06d}.parquet")

    def done(self):
        """Indices of the slabs whose file is complete (a file is renamed into place only after it is fully written)."""
        out = set()
        for f in os.listdir(self.out_dir):
            if f.startswith("slab_") and f.endswith(".parquet"):
                out.add(int(f[5:-8]))
        return out

    # -- writing -----------------------------------------------------------------------------------------------------
    def submit(self, slab, rows, pred, emb, release=None):
        """rows: list of (item, sample, gene_id, tissue) per prediction; pred [n] / emb [n, D] numpy views of a pinned
        buffer; release() is called once the buffer has been consumed (the caller may then reuse it)."""
        if self._err is not None:
            raise self._err
        self._q.put((slab, rows, pred, emb, release))

    def _drain(self):
        import pyarrow.parquet as pq
        pa = self.pa
        while True:
            job = self._q.get()
            if job is None:
                self._q.task_done()
                return
            slab, rows, pred, emb, release = job
            try:
                n = len(rows)
                cols = [pa.array(np.full(n, slab, np.int32)), pa.array(np.asarray([r[0] for r in rows], np.int32)),
                        pa.array([r[1] for r in rows], pa.string()), pa.array([r[2] for r in rows], pa.string()),
                        pa.array(np.asarray([r[3] for r in rows], np.int16)), pa.array(np.asarray(pred, np.float32))]
                if self.embeddings:
                    flat = pa.array(np.ascontiguousarray(emb, np.float32).reshape(-1))
                    cols.append(pa.FixedSizeListArray.from_arrays(flat, self.emb_dim))
                tmp = self.path(slab) + ".tmp"
                pq.write_table(pa.Table.from_arrays(cols, schema=self.schema), tmp, compression="zstd")
                os.replace(tmp, self.path(slab))
            except Exception as e:                       # noqa: BLE001  (surfaced on the caller's thread)
                self._err = e
            finally:
                if release is not None:
                    release()
                self._q.task_done()

    def flush(self):
        """Block until every submitted slab is on disk (or failed: the error is raised here)."""
        self._q.join()
        if self._err is not None:
            raise self._err

    def close(self):
        self._q.put(None)
        self._thread.join()
        if self._err is not None:
            raise self._err

    def read_all(self):
        """Every completed slab as one pandas frame, in slab / item order."""
        import pyarrow.parquet as pq
        tabs = [pq.read_table(self.path(s)) for s in sorted(self.done())]
        return self.pa.concat_tables(tabs).to_pandas() if tabs else None


class PinnedRing:
    """Two pinned host buffers per result tensor; a buffer is handed out again only after its consumer released it."""

    def __init__(self, n_buffers=2):
        self.free = queue.Queue()
        self.bufs = [{"pred": None, "emb": None} for _ in range(n_buffers)]
        for b in self.bufs:
            self.free.put(b)

    def acquire(self, n, d):
        b = self.free.get()
        if b["pred"] is None or b["pred"].numel() < n:
            b["pred"] = torch.empty(max(n, 1), dtype=torch.float32).pin_memory()
        if b["emb"] is None or b["emb"].numel() < n * d:
            b["emb"] = torch.empty(max(n * d, 1), dtype=torch.float32).pin_memory()
        return b

    def release(self, b):
        self.free.put(b)
