"""The five BASELINE.json configs through the CLI binary, byte for byte against the unmodified reference.

Inputs are regenerated from fixed seeds (ntsm_b200/synth_np.py); what the reference binary printed for
exactly these inputs was recorded in this repo's build container by `tools/config_parity.py
--make-expected` (sha256 / size / tallies per config in tests/golden/configs.json -- the reference
itself does not exist on the GPU box).  Sizes: cfg1 in full (1 M x 150 bp); cfg2 4 M reads in 16
files; cfg3 0.3 Gbases of ONT-like reads; cfg4 4 M paired-end reads as 8 .fq.gz, with and without
-m 10; cfg5 the full 10^6-site panel (26 M k-mers: the tables spill past L2, the pair table is
left unfolded and the k-mer bitmap is 2^30 bits -- the branch of ntsm_load_sites that the human panel
never takes).  Reference semantics: src/FingerPrint.hpp:46-103,270-311,473-564.
"""
import importlib.util
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("config_parity", os.path.join(ROOT, "tools", "config_parity.py"))
cp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cp)

EXPECTED = json.load(open(cp.EXPECTED))
CONFIGS = ["cfg1", "cfg2", "cfg3", "cfg4", "cfg4m", "cfg5"]


def test_expected_entries_cover_every_config():
    """CPU-side: the committed expectations hold all six runs at full test size, made by the reference."""
    full = EXPECTED["scale=1"]
    assert sorted(full) == sorted(CONFIGS)
    for cfg, e in full.items():
        assert e["rc"] == 0 and len(e["sha256"]) == 64 and e["stdout_bytes"] > 1000 and e["bases"] >= 150_000_000, cfg
    assert full["cfg4m"]["early_stop"] and full["cfg4m"]["bases"] < full["cfg4"]["bases"]


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("cfg"))


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS)
def test_config_counts_file_is_bit_exact(cfg, workdir):
    want = EXPECTED["scale=1"][cfg]
    sites, files, extra, threads = cp.make_inputs(cfg, workdir, 1.0)
    if not files[0].endswith(".gz"):         # (gzip's own output may differ between gzip builds; the inflated bytes may not)
        assert sum(os.path.getsize(f) for f in files) == want["input_bytes"], "generator drifted from the recorded inputs"
    env = dict(os.environ, NTSM_TIMING="1")
    info, out = cp.run(cp.OURS, sites, files, extra, threads, env=env)
    assert info["rc"] == 0, info.get("stderr_tail")
    assert info["stdout_bytes"] == want["stdout_bytes"]
    assert info["sha256"] == want["sha256"]
    assert info["bases"] == want["bases"] and info["kmers"] == want["kmers"] and info["hits"] == want["hits"]
    assert info["early_stop"] == want["early_stop"]
    if cfg == "cfg5":
        # the large-panel branch really ran: more than 5 M live k-mers
        assert info["hits"] > 0 and os.path.getsize(sites) > 400_000_000
