"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box; run with `gpurun --gpus 2|8 -- python -m pytest tests/test_gpu_multirank.py -m gpu`,
logs under profiles/): the multi-GPU paths of the C ABI against the 1-GPU result AND the oracle on the same reads.

  * one process per GPU (torch.distributed.run -> tests/multirank_worker.py): mhapb_comm_init_rank + mhapb_dist_search_self /
    mhapb_dist_search_query_reads -- job-wide counters and the union of the per-rank hit sets;
  * one process, all GPUs: mhapb_multi_* (what a single JVM binds) and `mhap-b200 --devices`.
The reference analogue is the single-JVM result the partitions must reproduce (docs/source/quickstart.rst:23).
"""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from mhap_b200 import native, synth
from mhap_b200.distributed import hits_digest, sorted_hits
from tests.gpu_common import assert_same_hits_bulk, engine
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _same(got, exp):
    assert len(got) == len(exp), (len(got), len(exp))
    g, e = sorted_hits(got), sorted_hits(exp)
    for f in g.dtype.names:
        if f != "pad_":
            assert (g[f] == e[f]).all(), f
    assert hits_digest(got) == hits_digest(exp)


N_TOTAL, L, NQ = 6001, 4000, 2003        # odd totals: ragged shards on every world size


def _single_gpu_reference():
    g = synth.genome(11, max(L + 1, N_TOTAL * L // 12))
    bases, offs = synth.reads(g, 77, 0, N_TOTAL, L, 0.12)
    qb, qo = synth.reads(g, 78, 0, NQ, L, 0.12)
    qids = np.arange(1, NQ + 1, dtype=np.int64) + N_TOTAL
    p = native.SketchParams(16, 256, 12, 1536, 0, 116)
    sp = native.SearchParams(3, 0, 0.2, 0.78, 1, 0, 0, -1)
    e = engine()
    e.store_reset(p)
    e.store_add_reads(bases, offs)
    hs, ss = e.search_self(sp)
    hq, sq = e.search_query_reads(sp, qb, qo, qids)
    return (bases, offs, qb, qo, qids, p, sp), (hs, ss, hq, sq)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_ranks_reproduce_single_gpu_and_oracle(tmp_path):
    world = _n_gpus()
    (bases, offs, qb, qo, qids, p, sp), (hs, ss, hq, sq) = _single_gpu_reference()
    assert len(hs) > 1000 and len(hq) > 300
    # 1-GPU result == oracle (so that the N-rank comparison below is against the oracle too)
    st = orc.Store(num_hashes=256)
    st.add_reads(bases, offs, threads=os.cpu_count())
    qs = orc.Store(num_hashes=256)
    qs.add_reads(qb, qo, ids=qids, both_strands=False, threads=os.cpu_count())
    rs, rq = st.search_self(threads=os.cpu_count(), keep_all=True), st.search_query(qs, threads=os.cpu_count(), keep_all=True)
    assert_same_hits_bulk(hs, rs.hits, ss, rs.stats)
    assert_same_hits_bulk(hq, rq.hits, sq, rq.stats)

    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multirank_worker.py"), str(tmp_path), str(N_TOTAL), str(L), str(NQ)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    self_parts = [np.load(tmp_path / f"self_{k}.npy") for k in range(world)]
    query_parts = [np.load(tmp_path / f"query_{k}.npy") for k in range(world)]
    stats = [json.load(open(tmp_path / f"stats_{k}.json")) for k in range(world)]
    for k in range(world):
        assert stats[k]["self"] == ss, (k, stats[k]["self"], ss)            # job-wide counters on every rank
        assert stats[k]["query"] == sq, (k, stats[k]["query"], sq)
    assert sum(s["n_store"] for s in stats) == 2 * N_TOTAL
    _same(np.concatenate(self_parts), hs)
    _same(np.concatenate(query_parts), hq)
    # a pair is reported by the rank that stores its target: the per-rank sets are disjoint by target id
    owner = {}
    for k, part in enumerate(self_parts):
        for t in np.unique(part["to_id"]):
            assert owner.setdefault(int(t), k) == k
    print(f"multirank parity ok: world={world} self_hits={len(hs)} query_hits={len(hq)} digest={hits_digest(hs)} counters={ss} "
          f"gather_ms={[round(s['gather_ms'], 2) for s in stats]} nccl={stats[0]['nccl']}")


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_multi_engine_one_process_all_gpus():
    world = _n_gpus()
    (bases, offs, qb, qo, qids, p, sp), (hs, ss, hq, sq) = _single_gpu_reference()
    m = native.MultiEngine(list(range(world)))
    m.store_reset(p)
    # reads arrive in batches, as from the streaming FASTA producer
    cuts = [0, 1500, 1501, 4000, N_TOTAL]
    for a, b in zip(cuts[:-1], cuts[1:]):
        m.store_add_reads(bases, offs[a:b + 1], np.arange(a + 1, b + 1, dtype=np.int64))
    assert m.store_size() == 2 * N_TOTAL
    h, s = m.search_self(sp)
    assert s == ss
    _same(h, hs)
    h2, s2 = m.search_query_reads(sp, qb, qo, qids)
    assert s2 == sq
    _same(h2, hq)
    m.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_cli_devices_output_equals_single_gpu(tmp_path):
    world = _n_gpus()
    g = synth.genome(5, 400_000)
    bases, offs = synth.reads(g, 9, 0, 1200, 3000, 0.1)
    fa = tmp_path / "reads.fasta"
    with open(fa, "w") as f:
        for i in range(1200):
            f.write(f">r{i}\n{bytes(bases[int(offs[i]):int(offs[i + 1])]).decode()}\n")
    exe = os.path.join(ROOT, "mhap_b200", "mhap-b200")
    one = subprocess.run([exe, "-s", str(fa), "--num-hashes", "256"], capture_output=True, text=True, timeout=600)
    many = subprocess.run([exe, "-s", str(fa), "--num-hashes", "256", "--devices", ",".join(map(str, range(world)))], capture_output=True, text=True, timeout=600)
    assert one.returncode == 0 and many.returncode == 0, many.stderr[-2000:]
    assert sorted(one.stdout.splitlines()) == sorted(many.stdout.splitlines()) and len(one.stdout.splitlines()) > 100
    pick = lambda e: [l for l in e.splitlines() if l.startswith(("Total matches", "Average"))]
    assert pick(one.stderr) == pick(many.stderr)
