"""Input pipeline of the inference path: the reference's ``DataLoader.load_test_batch_flow``.

The reference builds a ``tf.data`` pipeline (reference ``data_loader.py:241-325``): per sample it decodes
``<seq>/<id>.jpg`` (the three frames side by side, ``tf.image.decode_jpeg`` on ``/cpu:0``), ``np.load``s
``<id>-flownet2.npy`` ``(4,H,W,2)``, ``<id>-seglabel.npy`` ``(3,H,W,1)`` and the depth file through ``tf.py_func``
with ``num_parallel_calls=4``, zips them, ``batch(B)`` and ``prefetch(8 * B)``; ``test_kitti_pose.py:104-114`` takes
``inputs_batch[0..4]`` = image, pose, flow, depth, seglabel from the iterator.  This module keeps those names and
that tuple order, and replaces TensorFlow's runtime by worker threads (file reads, ``np.load`` and the JPEG decoder
release the GIL) that fill **pinned** host batches a few steps ahead of the consumer, so that
``DAVO.inference(inputs=...)`` finds its input ready and streams it to the GPU behind the previous batch's compute.

``decode='host'`` (default) decodes with PIL, i.e. libjpeg like TensorFlow's decoder; batches are pinned numpy arrays
for the host entry point.  ``decode='nvjpeg'`` hands the JPEG bytes to ``davo_decode_jpeg_batch`` (nvJPEG, on the GPU,
straight into the frame tensor) and returns CUDA tensors for the device entry point; its pixels are within a few levels
of libjpeg's, not identical (tests/test_gpu_parity.py measures the pose difference).
"""
from __future__ import annotations

import ctypes as C
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import parallel


def load_kitti_image_sequence_names(dataset_dir, frames, seq_length, load_pose=False, load_flow=False, load_depth=False,
                                    load_seglabel=False):
    """Reference ``test_kitti_pose.py:32-72``: the file lists of every valid target frame, with the reference's
    fallbacks (no seglabel -> the image names; no depth -> the SEGLABEL names; no flow -> the image names)."""
    names, tgt_inds, poses, flows, depths, segs = [], [], [], [], [], []
    for tgt_idx in range(len(frames)):
        if not parallel.is_valid_sample(frames, tgt_idx, seq_length):
            continue
        drive, fid = frames[tgt_idx].split(' ')
        base = os.path.join(dataset_dir, drive, fid)
        names.append(base + '.jpg')
        poses.append(base + '_cam.txt')
        flows.append(base + '-flownet2.npy')
        depths.append(base + '-monodepth2_depth.npy')
        segs.append(base + '-seglabel.npy')
        tgt_inds.append(tgt_idx)
    seglabel = segs if load_seglabel else names
    depth = depths if load_depth else segs
    flow = flows if load_flow else names
    return names, tgt_inds, (poses if load_pose else names), flow, depth, seglabel


class _Batches:
    """Iterator over batches; ``get_next()`` is the reference's name for ``__next__`` (test_kitti_pose.py:104)."""

    def __init__(self, loader, lists, system, decode, workers, prefetch, hold=1):
        self.loader, self.lists, self.system, self.decode = loader, lists, system, decode
        self._hold, self._held = max(1, hold), []          # batches the consumer may still be reading (asynchronous inference)
        self.n = len(lists[0])
        self.B = loader.batch_size
        self.n_batches = -(-self.n // self.B)
        self._q = queue.Queue(maxsize=max(1, prefetch))
        self._free = queue.Queue()
        self._stop = threading.Event()
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self._err = None
        self._buffers(prefetch + 1 + self._hold)
        self._thread = threading.Thread(target=self._produce, daemon=True)
        self._thread.start()
        self._served = 0

    # pinned staging: a batch's arrays are reused once the consumer has asked for a later batch
    def _buffers(self, count):
        import torch
        H, W, B = self.loader.img_height, self.loader.img_width, self.B
        pin = torch.cuda.is_available()

        def host(shape, dtype):
            t = torch.empty(shape, dtype=dtype)
            return (t.pin_memory() if pin else t).numpy()

        for _ in range(count):
            buf = {"flow": host((B, 4, H, W, 2), torch.float32), "seg": host((B, 3, H, W, 1), torch.float32)}
            if self.decode == "host":
                buf["img"] = host((B, H, 3 * W, 3), torch.uint8)
            if self.loader.read_depth:
                buf["depth"] = host((B, 3, H, W, 1), torch.float32)
            self._free.put(buf)

    def _load_sample(self, i):
        names, _, flows, depths, segs = self.lists
        H, W = self.loader.img_height, self.loader.img_width
        if self.decode == "host":
            from PIL import Image
            img = np.asarray(Image.open(names[i]).convert('RGB'), np.uint8)
            if img.shape != (H, 3 * W, 3):
                raise ValueError("%s is %s, expected %s" % (names[i], img.shape, (H, 3 * W, 3)))
        else:
            with open(names[i], 'rb') as f:
                img = f.read()
        flow = np.load(flows[i], allow_pickle=True) if self.loader.read_flow else None           # data_loader.py:280-282
        seg = np.load(segs[i], allow_pickle=True) if self.loader.read_seglabel else None
        depth = np.load(depths[i], allow_pickle=True) if self.loader.read_depth else None
        return img, flow, seg, depth

    def _produce(self):
        try:
            H, W = self.loader.img_height, self.loader.img_width
            for b in range(self.n_batches):
                idx = list(range(b * self.B, min((b + 1) * self.B, self.n)))
                futs = [self._pool.submit(self._load_sample, i) for i in idx]
                buf = self._free.get()
                if self._stop.is_set():
                    return
                jpegs = []
                for k, f in enumerate(futs):
                    img, flow, seg, depth = f.result()
                    if self.decode == "host":
                        buf["img"][k] = img
                    else:
                        jpegs.append(img)
                    if flow is not None:
                        buf["flow"][k] = np.asarray(flow, np.float32).reshape(4, H, W, 2)
                    if seg is not None:
                        buf["seg"][k] = np.asarray(seg, np.float32).reshape(3, H, W, 1)
                    if depth is not None:
                        buf["depth"][k] = np.asarray(depth, np.float32).reshape(3, H, W, 1)
                self._q.put((len(idx), buf, jpegs))
            self._q.put(None)
        except BaseException as e:  # noqa: BLE001  (surfaced to the consumer)
            self._err = e
            self._q.put(None)

    def __iter__(self):
        return self

    def __len__(self):
        return self.n_batches

    def __next__(self):
        while len(self._held) >= self._hold:
            self._free.put(self._held.pop(0))                # that batch has been consumed: its pinned arrays may be refilled
        item = self._q.get()
        if item is None:
            self._pool.shutdown(wait=False)
            if self._err is not None:
                raise self._err
            raise StopIteration
        n, buf, jpegs = item
        self._held.append(buf)
        self._served += 1
        depth = buf["depth"][:n] if "depth" in buf else None
        if self.decode == "host":
            return buf["img"][:n], None, buf["flow"][:n], depth, buf["seg"][:n]
        return self._to_device(n, buf, jpegs, depth)

    get_next = __next__

    def _to_device(self, n, buf, jpegs, depth):
        """nvJPEG decode into the frame tensor + upload of the planes the graph reads."""
        import torch
        sysm = self.system
        dev = "cuda:%d" % sysm.device
        H, W = self.loader.img_height, self.loader.img_width
        img = torch.empty((n, H, 3 * W, 3), dtype=torch.uint8, device=dev)
        ptrs = (C.c_void_p * n)(*[C.cast(C.c_char_p(j), C.c_void_p) for j in jpegs])
        sizes = (C.c_int64 * n)(*[len(j) for j in jpegs])
        stream = torch.cuda.current_stream(sysm.device).cuda_stream
        sysm._check(sysm._lib.davo_decode_jpeg_batch(sysm._h, ptrs, sizes, n, C.c_void_p(img.data_ptr()), C.c_void_p(stream)),
                    "davo_decode_jpeg_batch")
        flow = torch.zeros((n, 4, H, W, 2), dtype=torch.float32, device=dev)
        flow[:, :2].copy_(torch.from_numpy(buf["flow"][:n, :2]), non_blocking=True)        # planes 2, 3 are never read (davo.py:983-987)
        seg = torch.from_numpy(buf["seg"][:n]).to(dev, non_blocking=True)
        d = None if depth is None else torch.from_numpy(depth).to(dev, non_blocking=True)
        torch.cuda.current_stream(sysm.device).synchronize()       # the pinned batch may be refilled after the next call
        return img, None, flow, d, seg

    def close(self):
        self._stop.set()
        try:
            while True:
                self._q.get_nowait()
        except queue.Empty:
            pass
        self._free.put({})
        self._pool.shutdown(wait=False)


class DataLoader(object):
    """Reference ``data_loader.py:7-32`` (constructor arguments kept) with the inference entry point only."""

    def __init__(self, dataset_dir=None, batch_size=None, img_height=None, img_width=None, num_source=None,
                 num_scales=None, read_pose=False, read_flow=False, read_depth=False, read_seglabel=False,
                 data_aug=False, data_flip=False):
        self.dataset_dir, self.batch_size = dataset_dir, batch_size
        self.img_height, self.img_width, self.num_source = img_height, img_width, num_source
        self.read_pose, self.read_flow, self.read_depth, self.read_seglabel = read_pose, read_flow, read_depth, read_seglabel

    def load_test_batch_flow(self, image_sequence_names, image_sequence_poses, image_sequence_flows,
                             image_sequence_depths, image_sequence_seglabels, system=None, decode="host", workers=4,
                             prefetch=8, hold=1):
        """Reference ``data_loader.py:241-325``: an iterator over ``(image uint8 [B,H,3W,3], pose (None: never read by
        the graph), flow [B,4,H,W,2], depth [B,3,H,W,1] | None, seglabel [B,3,H,W,1])``.  ``workers`` = the reference's
        ``num_parallel_calls=4``; ``prefetch`` batches are kept ready (the reference: ``prefetch(batch_size * 8)``).
        ``decode='nvjpeg'`` needs ``system`` (a ``DAVO`` after ``setup_inference``) and yields CUDA tensors.  The arrays of
        a batch stay valid until ``hold`` further batches have been asked for (``hold=3`` for a consumer that keeps two
        ``inference_async`` calls in flight)."""
        if decode not in ("host", "nvjpeg"):
            raise ValueError("decode must be 'host' or 'nvjpeg'")
        if decode == "nvjpeg" and system is None:
            raise ValueError("decode='nvjpeg' decodes on the GPU of a DAVO handle: pass system=")
        lists = (list(image_sequence_names), list(image_sequence_poses), list(image_sequence_flows),
                 list(image_sequence_depths), list(image_sequence_seglabels))
        return _Batches(self, lists, system, decode, workers, prefetch, hold)

    def batch_unpack_image_sequence(self, image_seq, img_height, img_width, num_source):
        """Reference ``data_loader.py:537-557`` on a numpy / torch array ``[B,H,3W,C]``: (tgt, src stack on channels).
        The kernels do this inside ``pack8_kernel``; kept for callers that want the frames."""
        assert num_source == 2
        tgt = image_seq[:, :, img_width:2 * img_width]
        src0, src1 = image_seq[:, :, :img_width], image_seq[:, :, 2 * img_width:3 * img_width]
        cat = np.concatenate if isinstance(image_seq, np.ndarray) else __import__("torch").cat
        return tgt, cat([src0, src1], 3)
