"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sample-range split and the
frame reduce. The render itself is replaced by a deterministic per-sample fill."""
import os
import subprocess
import sys
import textwrap

from a_dive_into_ray_tracing_b200.dist import row_range, sample_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_range_partitions_exactly():
    for spp in (1, 7, 10, 500, 5000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                b, c = sample_range(spp, r, world)
                seen += list(range(b, b + c))
            assert seen == list(range(spp))


def test_row_range_partitions_exactly_on_tile_rows():
    for H in (2, 54, 225, 800, 2160):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                y0, y1 = row_range(H, r, world)
                assert y0 % 4 == 0 and (y1 % 4 == 0 or y1 == H) and 0 <= y0 <= y1 <= H
                rows += list(range(y0, y1))
            assert rows == list(range(H))


WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from a_dive_into_ray_tracing_b200.dist import sample_range, reduce_frames
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                            rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    H, W, spp = 6, 8, 37
    begin, count = sample_range(spp, rank, world)
    acc = torch.zeros(H, W, 4)
    for s in range(begin, begin + count):          # stand-in for rt_render_device
        acc[..., :3] += float(s + 1)
        acc[..., 3] += 1.0
    reduce_frames(acc)
    if rank == 0:
        assert torch.all(acc[..., 3] == spp), acc[..., 3]
        assert torch.all(acc[..., 0] == spp * (spp + 1) / 2)
    # image-space split: every rank fills its band of rows with all samples
    from a_dive_into_ray_tracing_b200.dist import row_range
    y0, y1 = row_range(H, rank, world)
    img = torch.zeros(H, W, 4)
    img[y0:y1, :, :3] = spp * (spp + 1) / 2        # stand-in for rt_render_rows_device
    img[y0:y1, :, 3] = spp
    reduce_frames(img)
    if rank == 0:
        assert torch.equal(img, acc)
        print("OK")
    dist.destroy_process_group()
""") % ROOT


def test_two_rank_reduce_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD="2", PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK" in outs[0][0]
