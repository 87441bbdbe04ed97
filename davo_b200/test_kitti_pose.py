"""Pose inference CLI with the reference's flags and output file.

Mirrors reference ``test_kitti_pose.py``: same flags (``:20-29``), same loop --
batches of ``batch_size`` samples through ``DAVO.inference(sess, mode='pose')``
(``:133-135``), zero target pose inserted / ``pose_vec2mat`` / first sample's
tgt->src0 then every sample's inv(tgt->src1) (``:136-145``), chained in fp64
(``:147-149``) and written as ``<output_dir>/<seq>-pred_kitti_pose.txt`` with 12
floats per line (``:116, 150-153``).

Input sources: ``--synthetic N`` generates an N-frame seeded stream (no dataset is
available offline); otherwise ``--concat_img_dir`` must hold the reference's dump
(``<seq>/<id>.jpg`` decoded by PIL if present, ``<id>-flownet2.npy``,
``<id>-seglabel.npy``; reference ``:45-49``).  ``--ckpt_file`` is an ``.npz`` of
``{tf variable name: array}``; with ``--synthetic`` and no checkpoint, TF-default
random init (seed 8964) is used.

Multi-GPU: launch with ``python -m torch.distributed.run --nproc-per-node N``; samples
are sharded contiguously by rank, poses all-gathered over NCCL, rank 0 composes and writes.
"""
from __future__ import annotations

import argparse
import os
from glob import glob

import numpy as np

from . import geo_utils, parallel, synthetic
from .davo import DAVO


def build_parser():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--batch_size", type=int, default=1, help="The size of of a sample batch")
    ap.add_argument("--img_height", type=int, default=128, help="Image height")
    ap.add_argument("--img_width", type=int, default=416, help="Image width")
    ap.add_argument("--seq_length", type=int, default=3, help="Sequence length for each example")
    ap.add_argument("--test_seq", type=int, default=9, help="Sequence id to test")
    ap.add_argument("--concat_img_dir", type=str, default=None, help="Preprocess image dataset directory")
    ap.add_argument("--output_dir", type=str, default=None, help="Output directory")
    ap.add_argument("--ckpt_file", type=str, default=None,
                    help="checkpoint: the prefix of a TensorFlow checkpoint (model-<step>) or an .npz of TF variables")
    ap.add_argument("--version", type=str, default="v1", help="version")
    ap.add_argument("--synthetic", type=int, default=0, help="use a seeded synthetic stream of this many frames")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--all_pairs", action="store_true",
                    help="compute both poses of every sample as the reference graph does; by default only "
                         "the poses the trajectory is composed from are computed (reference :143-145), "
                         "which writes the same file with half the work")
    ap.add_argument("--jpeg_decode", choices=["host", "nvjpeg"], default="host",
                    help="where the <id>.jpg frame triples of a dump are decoded: 'host' = PIL / libjpeg in the loader's "
                         "worker threads (TensorFlow's decoder family: the default), 'nvjpeg' = on the GPU, straight into "
                         "the frame tensor (pixels within a few levels of libjpeg's)")
    ap.add_argument("--loader_workers", type=int, default=4, help="reader / decoder threads (reference: num_parallel_calls=4)")
    ap.add_argument("--reference_batch_semantics", action="store_true",
                    help="at --batch_size > 1 write exactly the file the reference's loop writes: its `if i == 0` "
                         "(reference :143) tests the batch index, so EVERY sample of the first batch contributes its "
                         "tgt->src0 pose, and the duplicated samples that pad the last batch (reference :96-101) are "
                         "composed as well.  Default: the intended trajectory (first sample's tgt->src0 only, padding "
                         "trimmed), which is what the reference writes at its shipped --batch_size 1")
    return ap


class SyntheticStream:
    """Samples of a synthetic N-frame sequence, generated in seeded blocks of 64."""

    def __init__(self, n_frames, h, w, seed, depth_from="none"):
        self.n = n_frames - 2
        self.h, self.w, self.seed, self.depth_from = h, w, seed, depth_from
        self._blk, self._data = None, None

    def sample(self, i):
        blk = i // 64
        if blk != self._blk:
            self._data = synthetic.make_inputs(64, self.h, self.w, seed=self.seed + blk)
            if self.depth_from == "depth":
                self._data += (synthetic.make_depth(64, self.h, self.w, seed=4321 + self.seed + blk),)
            elif self.depth_from == "seglabel":
                self._data += (self._data[2],)
            self._blk = blk
        return tuple(a[i % 64] for a in self._data)


def depth_source(version):
    """Which file the reference's CLI feeds as ``input_depth`` (reference test_kitti_pose.py:48, 59-62, 91-94).

    ``read_depth = "depth" in version`` there -- but the GRAPH reads input_depth whenever "depth" OR "disp" is in
    the version (davo.py:960).  So a "disp"-only version (``-se_disp*``) is fed the SEGLABEL file as its depth
    (`depth = image_sequence_seglabels` when load_depth is False).  Mirrored, not repaired: "depth" ->
    ``<id>-monodepth2_depth.npy``, "seglabel" -> ``<id>-seglabel.npy``, "none" -> the graph does not read it.
    """
    if "depth" in version:
        return "depth"
    return "seglabel" if "disp" in version else "none"


class DumpStream:
    """The reference's on-disk dump (reference test_kitti_pose.py:33-72, doc/preprocessing.md)."""

    def __init__(self, root, seq, h, w, seq_length, depth_from="none"):
        d = os.path.join(root, '%.2d' % seq)
        half = int((seq_length - 1) / 2)
        n_frames = len(glob(d + '/*.jpg')) + 2 * half
        frames = ['%.2d %.6d' % (seq, n) for n in range(n_frames)]
        self.ids = [frames[i].split(' ')[1] for i in range(n_frames)
                    if parallel.is_valid_sample(frames, i, seq_length)]
        self.dir, self.n, self.h, self.w, self.depth_from = d, len(self.ids), h, w, depth_from

    def file_lists(self, depth_from="none"):
        """(names, poses, flows, depths, seglabels) as reference test_kitti_pose.py:32-72 builds them; the depth list is
        the seglabel list unless the version says "depth" (:59-62)."""
        b = [os.path.join(self.dir, fid) for fid in self.ids]
        segs = [x + '-seglabel.npy' for x in b]
        depths = [x + '-monodepth2_depth.npy' for x in b] if depth_from == "depth" else segs
        return [x + '.jpg' for x in b], [x + '_cam.txt' for x in b], [x + '-flownet2.npy' for x in b], depths, segs

    def sample(self, i):
        from PIL import Image
        fid = self.ids[i]
        img = np.asarray(Image.open(os.path.join(self.dir, fid + '.jpg')).convert('RGB'), np.uint8)
        flow = np.load(os.path.join(self.dir, fid + '-flownet2.npy')).astype(np.float32)
        seg = np.load(os.path.join(self.dir, fid + '-seglabel.npy')).astype(np.float32).reshape(3, self.h, self.w, 1)
        if self.depth_from == "none":
            return img, flow, seg
        if self.depth_from == "seglabel":
            return img, flow, seg, seg
        depth = np.load(os.path.join(self.dir, fid + '-monodepth2_depth.npy')).astype(np.float32)
        return img, flow, seg, depth.reshape(3, self.h, self.w, 1)


def main(argv=None):
    FLAGS = build_parser().parse_args(argv)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("test_kitti_pose: no CUDA device; the davo_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if FLAGS.output_dir and rank == 0 and not os.path.isdir(FLAGS.output_dir):
        os.makedirs(FLAGS.output_dir)
    H, W, B = FLAGS.img_height, FLAGS.img_width, FLAGS.batch_size
    if FLAGS.synthetic:
        stream = SyntheticStream(FLAGS.synthetic, H, W, FLAGS.seed, depth_source(FLAGS.version))
    else:
        if not FLAGS.concat_img_dir:
            raise SystemExit("test_kitti_pose: give --concat_img_dir or --synthetic N")
        stream = DumpStream(FLAGS.concat_img_dir, FLAGS.test_seq, H, W, FLAGS.seq_length, depth_source(FLAGS.version))
    n = stream.n
    order = list(range(n))
    if FLAGS.reference_batch_semantics:
        # the reference pads the SAMPLE LIST to a multiple of the batch size before anything else (:96-101) and
        # composes the duplicates; here that list is what gets sharded
        order = parallel.complete_batch_size(order, B)
        n = len(order)
    # shard, then fill the last batch by repeating the last item (reference common_utils.py:8-13)
    idx = [order[k] for k in parallel.padded_indices(n, rank, world)]
    first = parallel.padded_indices(n, rank, world)[0] if n else 0      # global position of this rank's first item
    n_local = len(idx)
    idx = parallel.complete_batch_size(list(idx), B)

    system = DAVO(version=FLAGS.version)
    system.setup_inference(H, W, "davo", FLAGS.seq_length, B, device=local)
    if FLAGS.ckpt_file:
        system.load_weights(FLAGS.ckpt_file)
    elif FLAGS.synthetic:
        system.load_weights(synthetic.init_weights(FLAGS.version))
    else:
        raise SystemExit("test_kitti_pose: --ckpt_file is required with a real dataset")

    if world > 1:
        system.init_comm(rank, world)                                  # the library's own NCCL communicator
    poses = torch.empty((len(idx), 2, 6), dtype=torch.float32, device="cuda:%d" % local)
    batches, in_flight = None, []
    if not FLAGS.synthetic:
        # the reference's input pipeline (test_kitti_pose.py:90-114): file lists -> DataLoader.load_test_batch_flow;
        # worker threads read, decode and fill pinned batches ahead of the loop below
        from .data_loader import DataLoader
        dsrc = depth_source(FLAGS.version)
        lists = stream.file_lists(dsrc)
        loader = DataLoader(FLAGS.concat_img_dir, B, H, W, FLAGS.seq_length - 1, read_flow=True, read_depth=dsrc != "none",
                            read_seglabel=True)
        batches = loader.load_test_batch_flow(*[[l[j] for j in idx] for l in lists], system=system,
                                              decode=FLAGS.jpeg_decode, workers=FLAGS.loader_workers, hold=3)
    for i in range(len(idx) // B):                                     # reference :133
        if batches is not None:
            img, _, flow, depth, seg = batches.get_next()              # reference :104-114: inputs_batch[0..4]
            inputs = (img, flow, seg) if depth is None else (img, flow, seg, depth)
        else:
            batch = [stream.sample(j) for j in idx[i * B:(i + 1) * B]]
            inputs = tuple(np.stack([s[k] for s in batch]) for k in range(len(batch[0])))   # (img, flow, seg[, depth])
        # reference :143-145 reads pose[s,1] of every sample and pose[0,0] of the sequence's first
        if FLAGS.all_pairs or (FLAGS.reference_batch_semantics and B > 1 and first + i * B < B):
            sel = 'all'                       # reference semantics: batch 0 contributes tgt->src0 of every sample
        else:
            sel = 'trajectory_first' if (i == 0 and first == 0) else 'trajectory'
        if isinstance(inputs[0], np.ndarray) and not system.config.batch_norm:
            # host arrays: queue this batch and collect the one before last, so that its copies run under the previous
            # batch's compute (reference :135 is a blocking sess.run behind tf.data's prefetch: the same overlap)
            in_flight.append((i, system.inference_async(inputs, pairs=sel)))
            while len(in_flight) > 2:
                j, h = in_flight.pop(0)
                poses[j * B:(j + 1) * B] = torch.as_tensor(h.result()['pose'])
        else:
            pred = system.inference(None, mode='pose', inputs=inputs, pairs=sel)   # reference :135
            poses[i * B:(i + 1) * B] = torch.as_tensor(pred['pose'])
    for j, h in in_flight:
        poses[j * B:(j + 1) * B] = torch.as_tensor(h.result()['pose'])
    all_poses = parallel.gather_poses(poses[:n_local].contiguous(), n, system).cpu().numpy()
    if rank == 0:
        traj = geo_utils.compose_trajectory(all_poses, B, FLAGS.reference_batch_semantics)   # reference :136-149
        if FLAGS.output_dir:
            out = os.path.join(FLAGS.output_dir, '%.2d-pred_kitti_pose.txt' % FLAGS.test_seq)
            if os.path.isfile(out):
                os.remove(out)
            geo_utils.write_kitti_trajectory(out, traj)
            print("Done. Please check %s" % out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return all_poses if rank == 0 else None


if __name__ == '__main__':
    main()
