"""Robustness of the host-side readers (prop.data, itoe.*, HNSW index files): random truncations and byte flips of valid
files must end in CDB_OK or a clean error, never in a crash or an endless loop.  Runs in a child process so that a memory
error cannot take the test session down."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent('''
    import os, shutil, sys
    import numpy as np
    sys.path.insert(0, sys.argv[1])
    import cosdata_b200 as cdb
    import oracle as orc
    from tests import mdgraph
    from tests.test_index_files import write_index_dir
    from tests.test_itoe_file import sample_store
    from tests.test_prop_file import write_prop_file

    tmp = sys.argv[2]
    rng = np.random.default_rng(1234)
    def mutate(path):
        b = bytearray(open(path, "rb").read())
        if not b:
            return
        kind = rng.integers(0, 3)
        if kind == 0:
            b = b[: int(rng.integers(0, len(b)))]
        elif kind == 1:
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        else:
            p = int(rng.integers(0, len(b)))
            b[p:p + 4] = bytes(rng.integers(0, 256, 4, dtype=np.uint8))
        open(path, "wb").write(bytes(b))
    def attempt(fn):
        try:
            fn()
            return 0
        except cdb.CosdataError:
            return 1
    outcomes = [0, 0]
    # prop.data
    vecs = orc.synth_matrix(5, 30, 20)
    for st in (0, 2, 4, 5):
        for it in range(40):
            p = os.path.join(tmp, "prop.data")
            write_prop_file(p, cdb.StorageType(st), vecs, np.arange(30, dtype=np.uint32))
            mutate(p)
            outcomes[attempt(lambda: cdb.prop_file_load(p))] += 1
            outcomes[attempt(lambda: cdb.prop_file_load_metadata(p))] += 1
    # itoe store
    for it in range(120):
        d = os.path.join(tmp, "coll")
        shutil.rmtree(d, ignore_errors=True)
        sample_store(d, dim=8, n=60)
        names = sorted(os.listdir(d))
        mutate(os.path.join(d, names[int(rng.integers(0, len(names)))]))
        outcomes[attempt(lambda: cdb.itoe_load(d))] += 1
        outcomes[attempt(lambda: cdb.itoe_get(d, int(rng.integers(0, 70000))))] += 1
    # HNSW index files
    vecs, mg = mdgraph.build(n=80, dim=8, levels=2, nb=4, nb0=8, storage_type=4, metric=0, seed=3)
    for it in range(120):
        d = os.path.join(tmp, "idx")
        shutil.rmtree(d, ignore_errors=True)
        rl, pl = write_index_dir(d, vecs, mg, seed=it)
        names = sorted(os.listdir(d))
        mutate(os.path.join(d, names[int(rng.integers(0, len(names)))]))
        def load():
            h = cdb.HnswFiles(d, rl, pl)
            h.close()
        outcomes[attempt(load)] += 1
    print("ok", outcomes[0], "errors", outcomes[1])
''')


def test_readers_survive_damaged_files(tmp_path):
    script = tmp_path / "fuzz_child.py"
    script.write_text(CHILD)
    r = subprocess.run([sys.executable, str(script), ROOT, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    words = r.stdout.split()
    assert words[0] == "ok" and int(words[3]) > 50 and int(words[1]) > 0      # both clean errors and surviving loads occurred
