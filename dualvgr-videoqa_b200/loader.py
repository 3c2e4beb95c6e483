"""Input pipeline for the train / eval engines — what the reference's DataLoader.py:61-84 does per sample
(open both HDF5 files, index one video, build fp32 tensors; 2.79 MB per sample over a pageable H2D copy, train.py:134),
restated for a step that takes ~5 ms:

  * FeatureStore: the appearance / motion features of all videos memory-mapped ONCE (no per-sample file open), optionally
    stored as bf16 (half the bytes on disk, in page cache and on the host link; the kernels take bf16 features directly,
    dvgr_prep_features_ex). Layout = the reference's HDF5 datasets (preprocess/preprocess_features.py:143-203):
    resnet_features [V, N, 16, 2048], resnext_features [V, N, 2048], ids [V]; read from .npy files (np.memmap) — or from
    the original .h5 files when h5py is importable.
  * PinnedBatchLoader: a background thread gathers the next batches into a ring of PINNED host buffers; the consumer issues
    the host-to-device copies on a copy stream (double-buffered device slots) and hands device tensors to
    TrainEngine.load_batch / EvalEngine.step. The training stream never waits on a page fault or a pageable copy."""
import queue
import threading

import numpy as np
import torch


class FeatureStore:
    def __init__(self, appearance, motion, ids=None):
        """appearance [V, N, F, Dv], motion [V, N, Dv]: numpy arrays / memmaps (float32, or uint16 holding bf16 bits)."""
        self.app, self.mot = appearance, motion
        self.ids = np.arange(appearance.shape[0]) if ids is None else np.asarray(ids)
        self.row = {int(v): i for i, v in enumerate(self.ids)}
        self.bf16 = appearance.dtype == np.uint16

    @classmethod
    def open(cls, appearance_path, motion_path):
        if appearance_path.endswith(".h5"):
            import h5py                                   # optional: the reference's own files
            fa, fm = h5py.File(appearance_path, "r"), h5py.File(motion_path, "r")
            return cls(fa["resnet_features"], fm["resnext_features"], fa["ids"][()])
        return cls(np.load(appearance_path, mmap_mode="r"), np.load(motion_path, mmap_mode="r"))

    @staticmethod
    def to_bf16_bits(x):
        """float32 array -> uint16 array of bf16 bit patterns (round to nearest even), the storage format of a bf16 store."""
        u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
        return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)

    def torch_dtype(self):
        return torch.bfloat16 if self.bf16 else torch.float32


class PinnedBatchLoader:
    """Iterates device batches (app, mot, question, question_len, answers[, extra...]) over `samples`.

    samples: dict of host arrays indexed by sample: 'video_idx' (row in the store, or 'video_id' resolved through store.ids),
    'question' [S, L] int64, 'question_len' [S] int64, 'answer' [S] int64, optional 'category' [S] int64."""

    def __init__(self, store, samples, batch_size, device, shuffle=False, seed=0, drop_last=True, slots=3):
        self.store, self.samples, self.B, self.dev = store, samples, batch_size, torch.device(device)
        self.shuffle, self.rng, self.drop_last, self.slots = shuffle, np.random.default_rng(seed), drop_last, max(2, slots)
        vid = samples.get("video_idx")
        if vid is None:
            vid = np.array([store.row[int(v)] for v in samples["video_id"]])
        self.vid = np.asarray(vid)
        self.S = len(self.vid)
        N, F, Dv = store.app.shape[1:]
        L = samples["question"].shape[1]
        fdt = store.torch_dtype()
        self.has_cat = "category" in samples

        def slot(pin):
            kw = dict(pin_memory=True) if pin else dict(device=self.dev)
            d = {"app": torch.empty((batch_size, N, F, Dv), dtype=fdt, **kw),
                 "mot": torch.empty((batch_size, N, Dv), dtype=fdt, **kw),
                 "q": torch.empty((batch_size, L), dtype=torch.int64, **kw),
                 "qlen": torch.empty((batch_size,), dtype=torch.int64, **kw),
                 "ans": torch.empty((batch_size,), dtype=torch.int64, **kw)}
            if self.has_cat:
                d["cat"] = torch.empty((batch_size,), dtype=torch.int64, **kw)
            return d

        self.host = [slot(True) for _ in range(self.slots)]
        self.device_slots = [slot(False) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.bytes_per_batch = sum(t.numel() * t.element_size() for t in self.host[0].values())

    def __len__(self):
        return self.S // self.B if self.drop_last else (self.S + self.B - 1) // self.B

    def _fill(self, slot, idx):
        """Gathers samples `idx` into pinned slot `slot` (runs on the worker thread; numpy releases the GIL in the copies)."""
        n = len(idx)
        order = np.argsort(self.vid[idx], kind="stable")            # ascending rows: sequential reads from the memmap
        rows = self.vid[idx][order]
        h = self.host[slot]
        app = h["app"].view(torch.uint16).numpy() if self.store.bf16 else h["app"].numpy()
        mot = h["mot"].view(torch.uint16).numpy() if self.store.bf16 else h["mot"].numpy()
        for k, r in zip(order, rows):
            app[k] = self.store.app[r]
            mot[k] = self.store.mot[r]
        h["q"][:n] = torch.from_numpy(np.ascontiguousarray(self.samples["question"][idx]))
        h["qlen"][:n] = torch.from_numpy(np.ascontiguousarray(self.samples["question_len"][idx]))
        h["ans"][:n] = torch.from_numpy(np.ascontiguousarray(self.samples["answer"][idx]))
        if self.has_cat:
            h["cat"][:n] = torch.from_numpy(np.ascontiguousarray(self.samples["category"][idx]))
        return n

    def __iter__(self):
        order = self.rng.permutation(self.S) if self.shuffle else np.arange(self.S)
        batches = [order[i:i + self.B] for i in range(0, self.S, self.B)]
        if self.drop_last and batches and len(batches[-1]) < self.B:
            batches.pop()
        free, ready = queue.Queue(), queue.Queue()
        for s in range(self.slots):
            free.put(s)

        def worker():
            try:
                for idx in batches:
                    s = free.get()
                    ready.put((s, self._fill(s, idx)))
                ready.put(None)
            except BaseException as e:                                # surface worker failures in the consumer
                ready.put(e)

        threading.Thread(target=worker, daemon=True).start()
        cur = torch.cuda.current_stream(self.dev)
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        for e in consumed:
            e.record(cur)
        pending = None                     # (device slot, n, copy-done event, pinned slot)
        k = 0

        def start_copy(item):
            nonlocal k
            s, n = item
            d = k % 2
            k += 1
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(consumed[d])            # the step that read this device slot has finished
                for name, t in self.host[s].items():
                    self.device_slots[d][name][:n].copy_(t[:n], non_blocking=True)
                done = torch.cuda.Event()
                done.record(self.copy_stream)
            return d, n, done, s

        while True:
            item = ready.get()
            if isinstance(item, BaseException):
                raise item
            nxt = start_copy(item) if item is not None else None
            if pending is not None:
                d, n, done, s = pending
                cur.wait_event(done)
                done.synchronize()                                   # the pinned slot may be refilled once its copy is done
                free.put(s)
                dev = self.device_slots[d]
                out = [dev["app"][:n], dev["mot"][:n], dev["q"][:n], dev["qlen"][:n], dev["ans"][:n]]
                if self.has_cat:
                    out.append(dev["cat"][:n])
                yield tuple(out)
                consumed[d].record(torch.cuda.current_stream(self.dev))
            pending = nxt
            if item is None:
                break
