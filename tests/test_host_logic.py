"""CPU tests of the host side: CLI flags / path munging mirror the reference, on-disk list parsing, test-mode summary
helpers, and the data-parallel contract (world_size-2 gloo): allreduce(sum) with 1/N folded into Adam == the reference's
get_average_grads (utils/utils.py:380-403) followed by apply_gradients."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_flags_match_reference_defaults():
    sys.path.insert(0, ROOT)
    import homography_CNN_synthetic as cli
    a = cli.build_parser().parse_args([])
    # defaults of code/homography_CNN_synthetic.py:49-85
    assert (a.mode, a.loss_type, a.num_gpus, a.batch_size, a.lr, a.min_lr) == ('train', 'l1_loss', 2, 128, 1e-4, .9e-4)
    assert (a.img_w, a.img_h, a.patch_size, a.do_augment, a.augment_list) == (320, 240, 128, 0.5, ['normalize'])
    assert a.use_batch_norm is False and a.resume is False and a.retrain is False and a.save_visual is True and a.visual is False
    a = cli.build_parser().parse_args(['--mode', 'test', '--loss_type', 'h_loss', '--data_path', '/d/', '--resume', 'True'])
    a = cli.resolve_paths(a)
    assert a.pts1_file == '/d/pts1.txt' and a.test_gt_file == '/d/test_gt.txt' and a.filenames_file == '/d/train_synthetic.txt'
    assert a.model_dir.endswith('synthetic_models/h_loss_normalize')
    assert a.log_dir.endswith('h_loss_normalizetest/h_loss_normalize/h_loss_normalizetest/')     # suffix applied twice (:94-100)
    assert a.resume is True
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(['--loss_type', 'nope'])
    assert cli.find_percentile([1, 2, 3, 4, 5, 6, 7, 8, 9, 10]) == [2.0, 5.0, 8.5]


def test_on_disk_list_parsing(tmp_path):
    from unsuperviseddeephomographyral2018_b200 import dataloader as dl
    (tmp_path / "list.txt").write_text("a.jpg a.jpg\nb.jpg b.jpg\n")
    np.savetxt(tmp_path / "pts1.txt", np.array([[50, 60, 178, 60, 178, 188, 50, 188], [45, 45, 173, 45, 173, 173, 45, 173]], dtype=float))
    np.savetxt(tmp_path / "gt.txt", np.array([[1, -2, 3, -4, 5, -6, 7, -8], [0] * 8], dtype=float))
    names, pts1, gt = dl.read_img_and_gt(str(tmp_path / "list.txt"), str(tmp_path / "pts1.txt"), str(tmp_path / "gt.txt"))
    assert names == [["a.jpg", "a.jpg"], ["b.jpg", "b.jpg"]] and pts1.shape == (2, 8) and gt[0, 1] == -2
    assert dl.count_text_lines(str(tmp_path / "list.txt")) == 2
    names, pts1, gt = dl.read_img_and_gt(str(tmp_path / "list.txt"), str(tmp_path / "pts1.txt"), None)
    assert gt is None
    assert dl.dataloader_params._fields == ('data_path', 'filenames_file', 'pts1_file', 'gt_file', 'mode', 'batch_size', 'img_h',
                                            'img_w', 'patch_size', 'augment_list', 'do_augment')


def test_model_params_namedtuple_matches_reference():
    from unsuperviseddeephomographyral2018_b200 import homography_model as hm
    assert hm.homography_model_params._fields == ('mode', 'batch_size', 'patch_size', 'img_w', 'img_h', 'loss_type', 'use_batch_norm',
                                                  'augment_list', 'leftright_consistent_weight')


def test_data_parallel_contract_gloo_world2(tmp_path):
    script = tmp_path / "dp.py"
    script.write_text(textwrap.dedent('''
        import os, sys
        sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        from oracle import oracle as O
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        torch.manual_seed(0)
        p = torch.randn(1000); m = torch.zeros(1000); v = torch.zeros(1000)
        tower = [torch.randn(1000, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        # reference: mean over towers, then Adam
        ref_p, _, _ = O.adam_step(p, O.average_grads([[t] for t in tower])[0], m, v, 1, 1e-3)
        # ours: every rank holds its own gradient, allreduce(sum), 1/N folded into the update (grad_scale)
        g = tower[rank].clone()
        dist.all_reduce(g)
        mine, _, _ = O.adam_step(p, g * (1.0 / world), m, v, 1, 1e-3)
        assert torch.allclose(mine, ref_p, atol=1e-7), (mine - ref_p).abs().max()
        # replicated update: all ranks end with identical parameters
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        assert all(torch.equal(gathered[0], x) for x in gathered)
        # the ENGINE's own two-phase reduction (engine.allreduce_two_phase / dp_slices: head slice, conv backward, conv slice)
        # on the real flat layout: every element is summed exactly once, the hook between the phases runs, and the per-rank
        # dropout seeds of the engine differ
        from unsuperviseddeephomographyral2018_b200 import engine as en, params as P
        specs = P.param_specs()
        n = P.total_floats(specs)
        head, convs = en.dp_slices(specs)
        cover = torch.zeros(n); cover[head] += 1; cover[convs] += 1
        assert bool((cover == 1).all()) and head.start == specs["model/fc1/fc1/weights"].offset
        flat = torch.full((n,), float(rank + 1))
        flat[specs["model/conv_block1/conv1/weights"].offset] = 10.0 * (rank + 1)
        seen = []
        en.allreduce_two_phase(flat, specs, None, between=lambda: seen.append(float(flat[-1])))
        tot = float(sum(range(1, world + 1)))
        assert seen == [tot], seen                                     # the head slice was already reduced when the hook ran
        assert float(flat[1]) == tot and float(flat[-1]) == tot and float(flat[specs["model/conv_block1/conv1/weights"].offset]) == 10.0 * tot
        # the sharded schedule of the multicast path (csrc/dp_update.cu): sum over ranks, Adam on this rank's shard only,
        # updated shards back to every replica == the replicated update (the shard table is engine.dp_shard_range)
        lo, hi, per = en.dp_shard_range(0, 1000, rank, world)
        gs = tower[rank].clone(); dist.all_reduce(gs)
        shard_p, _, _ = O.adam_step(p[lo:hi], gs[lo:hi] * (1.0 / world), m[lo:hi], v[lo:hi], 1, 1e-3)
        full_p = torch.zeros(per * world)
        pad = torch.zeros(per); pad[:hi - lo] = shard_p
        dist.all_gather(list(full_p.view(world, per).unbind(0)), pad)
        assert torch.equal(full_p[:1000], mine), (full_p[:1000] - mine).abs().max()
        # per-rank batch sharding as tf.split(batch, num_gpus) (homography_CNN_synthetic.py:199-207)
        full = torch.arange(8)
        assert torch.equal(full.chunk(world)[rank], full[rank * 4:(rank + 1) * 4])
        dist.destroy_process_group()
        print("rank", rank, "ok")
    ''' % ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_no_undefined_names_in_gpu_only_code():
    """Most of the package only runs on a GPU box; a forgotten import there costs a GPU round trip.  tools/undef_check.py
    flags names that are loaded but bound nowhere in their module."""
    import glob
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import undef_check
    files = (glob.glob(os.path.join(ROOT, "unsuperviseddeephomographyral2018_b200", "*.py")) + glob.glob(os.path.join(ROOT, "*.py")) +
             glob.glob(os.path.join(ROOT, "tests", "*.py")) + glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "oracle", "*.py")))
    assert len(files) > 30
    bad = [(os.path.relpath(f, ROOT), u) for f in files for u in undef_check.undefined_names(f)]
    assert not bad, bad


def test_dp_shard_table():
    """engine.dp_shard_range: 4-float aligned shards that cover the tensor exactly once, for even and ragged splits."""
    pytest.importorskip("torch")
    from unsuperviseddeephomographyral2018_b200 import engine as en, params as P
    w = P.param_specs()["model/fc1/fc1/weights"]
    cases = [(w.offset, int(np.prod(w.shape)), n) for n in (1, 2, 3, 4, 8)] + [(64, 1000, 3), (0, 4, 8), (128, 36, 5)]
    for begin, count, world in cases:
        cover = np.zeros(begin + count + 8, np.int32)
        sizes = []
        for r in range(world):
            lo, hi, per = en.dp_shard_range(begin, count, r, world)
            assert begin <= lo <= hi <= begin + count and lo % 4 == 0 and per % 4 == 0 and hi - lo <= per
            assert (hi - lo) % 4 == 0 or hi == begin + count
            cover[lo:hi] += 1
            sizes.append(hi - lo)
        assert (cover[begin:begin + count] == 1).all() and cover.sum() == count
        assert sizes == sorted(sizes, reverse=True)            # full shards first, the ragged / empty ones at the tail


def test_crc32c_native_and_numpy_agree():
    """udh_crc32c (C ABI, hardware crc32 / slicing-by-8) against the pure numpy CRC-32C and the standard check value."""
    from unsuperviseddeephomographyral2018_b200 import _lib, tf_checkpoint as tfc
    import ctypes
    assert tfc.crc32c_numpy(b"123456789") == 0xE3069283 and tfc.crc32c(b"123456789") == 0xE3069283
    rng = np.random.default_rng(3)
    for n in (1, 7, 8, 9, 63, 4097, 65535, 65536, (1 << 18) + 5):
        a = rng.integers(0, 256, n + 3, dtype=np.uint8)
        for off in (0, 3):                                     # aligned and unaligned starts
            b = a[off:off + n]
            nat = int(_lib.lib.udh_crc32c(ctypes.c_void_p(b.ctypes.data), b.size, 0))
            assert nat == tfc.crc32c_numpy(np.ascontiguousarray(b)) == tfc.crc32c(b), (n, off)
    # continuation: crc(a || b) from crc(a)
    a = rng.integers(0, 256, 1000, dtype=np.uint8)
    c1 = int(_lib.lib.udh_crc32c(ctypes.c_void_p(a.ctypes.data), 400, 0))
    c2 = int(_lib.lib.udh_crc32c(ctypes.c_void_p(a.ctypes.data + 400), 600, c1))
    assert c2 == tfc.crc32c_numpy(a)


def test_cli_background_checkpoint_writer(tmp_path):
    """homography_CNN_synthetic.save(): the state is snapshotted in the caller, the bundle is written by a thread; a second
    save joins the first, wait_for_save() leaves complete, readable bundles and an up-to-date `checkpoint` state file."""
    sys.path.insert(0, ROOT)
    import types
    import homography_CNN_synthetic as cli
    from unsuperviseddeephomographyral2018_b200 import params as P, tf_checkpoint as tfc
    specs = P.param_specs()
    n = P.total_floats(specs)

    class Eng(object):                                          # the two members save() touches
        def __init__(self):
            self.global_step = 0
            self.calls = 0

        def snapshot_tf_variables(self, with_optimizer=True):
            self.calls += 1
            flat = np.full(n, float(self.global_step), np.float32)
            return tfc.engine_state_to_variables(flat, None, None, self.global_step, specs)

    args = types.SimpleNamespace(model_dir=str(tmp_path), model_name="model.ckpt")
    eng = Eng()
    for step in (1000, 2000):
        eng.global_step = step
        cli.save(eng, args, step)                               # returns while the writer thread runs
    cli.wait_for_save()
    assert eng.calls == 2 and cli._save_thread is None
    assert tfc.latest_checkpoint(str(tmp_path)).endswith("model.ckpt-2000")
    for step in (1000, 2000):
        back = tfc.read_checkpoint(os.path.join(str(tmp_path), "model.ckpt-%d" % step))      # verifies every CRC
        assert int(np.asarray(back["Variable"]).reshape(-1)[0]) == step      # global_step is the unnamed `Variable` of the graph
        assert float(back["model/fc2/fc2/biases"][0]) == float(step)
    eng.global_step = 3000
    cli.save(eng, args, 3000, background=False)
    assert tfc.latest_checkpoint(str(tmp_path)).endswith("model.ckpt-3000") and cli._save_thread is None


def test_named_checkpoint_roundtrip(tmp_path):
    """TF-Slim variable names / shapes: export -> import is lossless, wrong shapes are rejected."""
    from unsuperviseddeephomographyral2018_b200 import params as P
    flat = P.init_flat(3)
    path = str(tmp_path / "ckpt.npz")
    P.save_named_npz(path, flat, extra={"global_step": np.array(1234)})
    z = np.load(path)
    assert z["model/conv_block1/conv1/weights"].shape == (3, 3, 2, 64) and z["model/fc1/fc1/weights"].shape == (32768, 1024)
    assert z["model/fc2/fc2/biases"].shape == (8,) and int(z["global_step"]) == 1234
    back = P.load_named_npz(path)
    assert np.array_equal(back, flat)
    bad = dict(z); bad["model/fc2/fc2/weights"] = np.zeros((8, 1024), np.float32)
    np.savez(str(tmp_path / "bad.npz"), **bad)
    with pytest.raises(ValueError):
        P.load_named_npz(str(tmp_path / "bad.npz"))


def test_tf_checkpoint_v2_roundtrip_and_format(tmp_path):
    """Dependency-free TensorFlow checkpoint V2 writer / reader: CRC-32C known answer, table framing
    (footer magic, block CRCs), the reference graph's variable names incl. Adam slots and global_step, the NHWC-flatten fc1
    layout, corruption detection, and the `checkpoint` state file of tf.train.Saver."""
    import struct
    from unsuperviseddeephomographyral2018_b200 import params as P, tf_checkpoint as tfc
    assert tfc.crc32c(b"123456789") == 0xE3069283                          # CRC-32C (Castagnoli) check value
    big = np.random.default_rng(0).integers(0, 256, size=200003, dtype=np.uint8)
    assert tfc.crc32c(big) == (tfc._crc_update_small(0xFFFFFFFF, big.tolist()) ^ 0xFFFFFFFF)    # lane-parallel path == serial path
    assert tfc.unmask_crc(tfc.mask_crc(0x12345678)) == 0x12345678
    specs = P.param_specs()
    n = P.total_floats(specs)
    rng = np.random.default_rng(1)
    flat = P.init_flat(5)
    # fc1 rows follow the NHWC flatten order (h*16 + w)*128 + c: tag one element per (h, w, c) corner case
    s1 = specs["model/fc1/fc1/weights"]
    fc1 = flat[s1.offset:s1.offset + s1.size].reshape(s1.shape)
    fc1[(3 * 16 + 5) * 128 + 7, 11] = 1234.5
    m = rng.normal(size=n).astype(np.float32) * 1e-3; v = rng.uniform(0, 1e-4, size=n).astype(np.float32)
    variables = tfc.engine_state_to_variables(flat, m, v, 4321, specs)
    assert variables["model/conv_block1/conv1/weights"].shape == (3, 3, 2, 64) and variables["model/fc1/fc1/weights"].shape == (32768, 1024)
    assert "model/fc2/fc2/biases/Adam_1" in variables and variables["Variable"].dtype == np.int32 and int(variables["Variable"]) == 4321
    assert abs(float(variables["beta1_power"]) - 0.9 ** 4322) < 1e-12
    prefix = str(tmp_path / "m" / "model.ckpt-4321")
    tfc.write_checkpoint(prefix, variables)
    idx = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", idx[-8:])[0] == 0xdb4775248b80fb57            # LevelDB table magic
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(a.nbytes for a in variables.values())
    back = tfc.read_checkpoint(prefix)
    assert set(back) == set(variables)
    for k in variables:
        assert back[k].dtype == variables[k].dtype and np.array_equal(back[k], variables[k]), k
    f2, m2, v2, step = tfc.variables_to_engine_state(back, specs, n)
    assert step == 4321 and np.array_equal(f2, flat) and np.array_equal(m2[s1.offset:s1.offset + s1.size], m[s1.offset:s1.offset + s1.size])
    assert back["model/fc1/fc1/weights"][(3 * 16 + 5) * 128 + 7, 11] == np.float32(1234.5)
    # a flipped byte in the data shard is caught by the per-tensor CRC; a truncated index by the footer check
    with open(prefix + ".data-00000-of-00001", "r+b") as f:
        f.seek(1000); b = f.read(1); f.seek(1000); f.write(bytes([b[0] ^ 0xFF]))
    with pytest.raises(ValueError):
        tfc.read_checkpoint(prefix)
    open(str(tmp_path / "bad.index"), "wb").write(idx[:-9])
    open(str(tmp_path / "bad.data-00000-of-00001"), "wb").write(b"")
    with pytest.raises(ValueError):
        tfc.read_checkpoint(str(tmp_path / "bad"))
    # inference-only checkpoint (no slots): parameters load, optimiser state is reported absent
    small = {k: a for k, a in variables.items() if "/Adam" not in k and "power" not in k}
    tfc.write_checkpoint(str(tmp_path / "inf"), small)
    f3, m3, v3, step3 = tfc.variables_to_engine_state(tfc.read_checkpoint(str(tmp_path / "inf")), specs, n)
    assert m3 is None and v3 is None and step3 == 4321 and np.array_equal(f3, flat)
    # Saver state file
    tfc.update_checkpoint_state(str(tmp_path / "m"), prefix)
    assert tfc.latest_checkpoint(str(tmp_path / "m")) == prefix


def test_real_data_correspondence_metric_and_flags():
    """Real-data path: the correspondence metric of code/homography_CNN_real.py:578-612 restated in NumPy
    (against hand-built cases), and the entry script's flags / directory rules (:64-124)."""
    import torch
    from oracle import oracle as O
    from unsuperviseddeephomographyral2018_b200 import real_metrics as rm, dataloader as dl
    rng = np.random.default_rng(0)
    # getPerspectiveTransform == the reference's DLT (h33 = 1) and maps the four points exactly
    src = np.array([[24., 7.], [152., 7.], [152., 135.], [24., 135.]]); dst = src + rng.uniform(-20, 20, size=(4, 2))
    H = rm.get_perspective_transform(src, dst)
    assert abs(H[2, 2] - 1) < 1e-15 and np.abs(rm.perspective_transform(src, H) - dst).max() < 1e-9
    assert np.abs(H - O.solve_dlt(torch.tensor(src.reshape(1, 8)), torch.tensor((dst - src).reshape(1, 8)))[0].numpy()).max() < 1e-9
    # a perfect prediction gives ~0 error; a zero prediction gives exactly the identity error (and is not a "failure")
    pts1 = np.array([[24., 7., 152., 7., 152., 135., 24., 135.]])
    gt = rng.integers(-7, 8, size=(1, 8)).astype(np.float64)
    r = 240.0 / 142.0
    Hf = rm.get_perspective_transform((pts1.reshape(4, 2) * r).astype(np.float32), ((pts1 + gt).reshape(4, 2) * r).astype(np.float32))
    c1 = np.stack([rng.uniform(40, 280, 4), rng.uniform(40, 200, 4)], 1)
    c2 = rm.perspective_transform(c1, np.linalg.inv(Hf))
    corr = np.concatenate([c1.reshape(-1), c2.reshape(-1)])[None] * 2.0
    (h, ident, failed), = rm.correspondence_errors(gt, pts1, corr)
    assert h < 1e-4 and not failed and ident > 0.5
    (h0, ident0, failed0), = rm.correspondence_errors(np.zeros((1, 8)), pts1, corr)
    assert abs(h0 - ident0) < 1e-9 and abs(ident0 - np.sqrt(np.mean((c1 - c2) ** 2))) < 1e-9 and not failed0
    (h1, ident1, failed1), = rm.correspondence_errors(-3 * gt, pts1, corr)       # a prediction worse than identity is bounded
    assert failed1 and h1 == ident1
    # flags and directory rules
    import homography_CNN_real as real
    a = real.resolve_paths(real.build_parser().parse_args(["--data_path", "/d/", "--save_model_dir", "/m/real/", "--load_model_dir", "/m/syn/"]))
    assert (a.img_h, a.img_w, a.full_img_h, a.full_img_w, a.patch_size) == (142, 190, 240, 320, 128)
    assert a.finetune is True and a.gt_file is None and a.loss_type == "l1_loss" and a.numeric == "bf16x3"
    assert a.load_model_dir == "/m/syn/l1_loss_normalize" and a.save_model_dir == "/m/real/l1_loss_normalize"
    assert a.test_gt_file == "/d/test_gt.txt" and a.filenames_file == "/d/train_real.txt"
    assert dl.extended_dataloader_params._fields[-2:] == ('full_img_h', 'full_img_w')


def test_tensorboard_event_file_roundtrip(tmp_path):
    """The scalars the reference logs every 1000 steps (code/homography_CNN_synthetic.py:285-293,356-358) as a TFRecord /
    Event-proto file written without TensorFlow: framing CRCs, file_version header, tags, steps, values."""
    from unsuperviseddeephomographyral2018_b200 import tb_events
    w = tb_events.SummaryWriter(str(tmp_path / "log"))
    tags = ["Losses/Learning_rate"] + ["Losses/Total_%s" % k for k in ("h_loss", "rec_loss", "ssim_loss", "l1_loss", "l1_smooth_loss", "ncc_loss")]
    w.add_scalars({t: 0.5 + i for i, t in enumerate(tags)}, 0)
    w.add_scalars({"Losses/Learning_rate": 1e-4}, 1000)
    w.close()
    raw = open(w.path, "rb").read()
    assert b"brain.Event:2" in raw[:64] and os.path.basename(w.path).startswith("events.out.tfevents.")
    ev = tb_events.read_events(w.path)
    assert [s for s, _ in ev] == [0, 1000] and list(ev[0][1]) == tags
    assert ev[0][1]["Losses/Total_ncc_loss"] == 6.5 and abs(ev[1][1]["Losses/Learning_rate"] - 1e-4) < 1e-10
