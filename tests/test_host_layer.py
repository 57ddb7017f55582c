"""CPU suite, part 2: the product's host layer and its C ABI (no GPU compute calls).

* libabpoa_b200.so loads and exports every function include/abpoa.h and include/abpoa_gpu.h declare;
* struct layouts seen by ctypes match the header as compiled by gcc;
* writers: consensus FASTA / RC-MSA text produced by the product's abpoa_output() (alignments
  injected from the oracle) hash to the md5 of the reference CLI's stdout for the same input;
* the product refuses to align without a GPU instead of falling back to anything.
"""
import ctypes as C
import hashlib
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from abpoa_b200 import capi
from abpoa_b200.aligner import PoaConfig, PoaSession
from cases import AFFINE
from helpers import INPUTS, read_fasta
from oracle_binding import oracle_align

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())


def declared_functions(header: Path) -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    text = re.sub(r"//.*", "", text)
    return sorted(set(re.findall(r"\b(abpoa_\w+)\s*\(", text)))


@pytest.mark.parametrize("header", ["abpoa.h", "abpoa_gpu.h"])
def test_exports_every_declared_symbol(product_lib, header):
    names = declared_functions(ROOT / "include" / header)
    assert names, header
    missing = [n for n in names if not hasattr(product_lib.dll, n)]
    assert not missing, f"{header}: not exported: {missing}"


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "abpoa.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(abpoa_para_t),offsetof(abpoa_para_t,incr_fn),offsetof(abpoa_para_t,min_freq),sizeof(abpoa_node_t),"
                   "sizeof(abpoa_graph_t),sizeof(abpoa_res_t),sizeof(abpoa_cons_t),sizeof(abpoa_simd_matrix_t));return 0;}\n")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    got = list(map(int, subprocess.run([str(exe)], capture_output=True, check=True, text=True).stdout.split()))
    p = capi.abpoa_para_t
    want = [C.sizeof(p), p.incr_fn.offset, p.min_freq.offset, C.sizeof(capi.abpoa_node_t), C.sizeof(capi.abpoa_graph_t),
            C.sizeof(capi.abpoa_res_t), C.sizeof(capi.abpoa_cons_t), C.sizeof(capi.abpoa_simd_matrix_t)]
    assert got == want


def _fasta_names(path):
    return [ln[1:].split()[0] for ln in Path(path).read_text().splitlines() if ln.startswith(">")]


@pytest.mark.parametrize("tag,cfgkw,fname", [
    ("seq.fa -O 4 -E 2", dict(**AFFINE), "seq.fa"),
    ("seq.fa -O 4 -E 2 -r1", dict(out_msa=True, out_cons=False, **AFFINE), "seq.fa"),
    ("seq.fa -O 4 -E 2 -r2", dict(out_msa=True, **AFFINE), "seq.fa"),
    ("seq.fa", dict(), "seq.fa"),
    ("test.fa", dict(), "test.fa"),
    ("heter.fa -r2", dict(out_msa=True), "heter.fa"),
])
def test_writers_match_reference_cli_md5(product_lib, tmp_path, tag, cfgkw, fname):
    reads = read_fasta(INPUTS / fname)
    names = _fasta_names(INPUTS / fname)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    out = tmp_path / "out.txt"
    with PoaSession(PoaConfig(**cfgkw), product_lib) as s:
        s.reset(max(len(r) for r in reads))
        abs_ = s.ab.contents.abs.contents
        for i, r in enumerate(reads):
            _, res = oracle_align(s, r)
            s.add(r, res, len(reads))
        # names as abpoa_msa() would have stored them
        for i, nm in enumerate(names):
            b = nm.encode()
            buf = capi.libc_realloc(None, len(b) + 1)
            C.memmove(buf, b + b"\0", len(b) + 1)
            abs_.name[i].s = C.cast(buf, C.c_char_p)
            abs_.name[i].l = len(b)
            abs_.name[i].m = len(b) + 1
        fp = libc.fopen(str(out).encode(), b"w")
        product_lib.abpoa_output(s.ab, s.abpt, fp)
        libc.fclose(fp)
    assert hashlib.md5(out.read_bytes()).hexdigest() == GOLDEN["cli_md5"][tag]


def test_product_has_no_cpu_fallback(tmp_path):
    """Without a CUDA device the alignment entry point must die loudly (exit != 0)."""
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from abpoa_b200.aligner import PoaSession, PoaConfig\n"
        "from abpoa_b200 import synth\n"
        "s = PoaSession(PoaConfig())\n"
        "s.run_reads(synth.make_group(1, 3, 50, 0.05))\n"
        "print('ALIGNED')\n" % (str(ROOT), str(ROOT / "tests")))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert p.returncode != 0 and "ALIGNED" not in p.stdout
    assert "no CUDA device" in p.stderr or "CUDA" in p.stderr


def test_batch_engine_has_no_cpu_fallback():
    """The batched entry point must refuse to start without a CUDA device, too."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from abpoa_b200.batch import BatchEngine\n"
        "e = BatchEngine(n_workers=2)\n"
        "print('STARTED')\n" % str(ROOT))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert p.returncode != 0 and "STARTED" not in p.stdout
    assert "no CUDA device" in p.stderr or "CUDA" in p.stderr


def test_handle_reuse_after_reset_with_high_degree_nodes(product_lib):
    """A handle that held a bushy graph (nodes with more than 4 edges, spilled out of the inline
    slots) must give the same result as a fresh handle after abpoa_reset()."""
    from abpoa_b200 import synth
    from helpers import run_group
    from oracle_binding import oracle_align
    cfg = PoaConfig(out_msa=True)
    bushy = synth.make_group(77, 14, 120, 0.30)          # 30 % error: many alternative branches
    plain = synth.make_group(78, 6, 300, 0.05)
    fresh = run_group(product_lib, cfg, plain, use_oracle=True)
    with PoaSession(cfg, product_lib) as s:
        for reads in (bushy, plain):
            s.reset(400)
            for r in reads:
                _, res = oracle_align(s, r)
                s.add(r, res, len(reads))
        g = s.ab.contents.abg.contents
        assert max(g.node[i].in_edge_n for i in range(g.node_n)) >= 1
        s.generate()
        assert all(np.array_equal(x, y) for x, y in zip(s.consensus(), fresh["cons"]))
        assert all(np.array_equal(x, y) for x, y in zip(s.msa_rows(), fresh["msa"]))


def test_bench_rank_cpu_shares_are_disjoint_and_complete():
    """bench.py gives every rank of a multi-GPU run its own slice of the host's cores."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    allowed = sorted(os.sched_getaffinity(0))
    for world in (1, 2, 4):
        shares = [mod.rank_cpu_share(r, world) for r in range(world)]
        flat = [c for sh in shares for c in sh]
        if len(allowed) >= world:
            assert sorted(flat) == allowed and len(set(flat)) == len(flat), (world, shares)
