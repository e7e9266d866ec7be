"""CPU: the evidence the documents cite exists and says what they say.  Every `profiles/...` path named in DESIGN.md /
README.md / INTEGRATION.md / include/*.h must be a committed file; the headline numbers quoted in DESIGN.md must be the
ones in the committed bench lines (a number in prose that no file backs is how a README starts advertising 1655)."""
import json
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _text(rel):
    with open(os.path.join(ROOT, rel), encoding="utf-8") as f:
        return f.read()


def _last_json_line(rel):
    with open(os.path.join(ROOT, rel)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_every_cited_profile_file_exists():
    docs = ["DESIGN.md", "README.md", "INTEGRATION.md", "include/gab200_rasterizer.h", "bench.py"]
    missing = []
    for d in docs:
        for m in re.finditer(r"profiles/(r0\d/)?[A-Za-z0-9_.\-]+\.(?:jsonl|json|csv|log|txt)", _text(d)):
            if not os.path.exists(os.path.join(ROOT, m.group(0))):
                missing.append((d, m.group(0)))
    # short forms inside a sentence that already named the directory: `bwd_variants_b.jsonl`, `fps_sweep.jsonl`, ...
    for m in re.finditer(r"`([A-Za-z0-9_\-]+\.(?:jsonl|json|csv))`", _text("DESIGN.md")):
        name = m.group(1)
        if name.startswith(("bench_", "bwd_", "fps_", "ncu_", "nvls_", "n2_", "train_", "multi_", "scan_", "e2e_", "tile_", "fwd_")):
            if not any(os.path.exists(os.path.join(ROOT, "profiles", sub, name)) for sub in ("r02", "r01", "")):
                missing.append(("DESIGN.md", name))
    assert not missing, f"documents cite evidence files that are not in the repository: {missing}"


def test_headline_numbers_in_design_are_the_committed_bench_line():
    line = _last_json_line("profiles/r02/bench_1gpu_final.json")
    design = _text("DESIGN.md")
    assert line["n_gpus"] == 1 and line["config"]["splats"] == 100_000 and line["config"]["width"] == 1920
    assert f"{round(line['value'])} frames/s" in design, "DESIGN.md quotes a headline value the committed line does not hold"
    assert f"{round(line['e2e']['value'])} frames/s" in design
    assert line["clocks"]["reasons"] == [] and line["graph_overflow"] is False
    assert line["gpu_launches"] > 0 and line["roofline"]["kernel"] == "blend_bwd"
    # the step really is the sum of its kernels: value within 5 % of the per-stage events
    assert abs(line["ms_per_step"] - sum(line["stage_ms"].values())) <= 0.05 * line["ms_per_step"]
    pc = line["parity_check"]
    assert pc["image_values_over_1e-4"] < 100 and all(v["beyond_atol_rtol"] <= 16 for v in pc["grads"].values())


def test_multi_gpu_lines_say_what_was_timed():
    for rel, n in (("profiles/r02/bench_n2_deferred_auto.json", 2), ("profiles/r02/bench_n8_auto.json", 8),
                   ("profiles/r02/bench_n8_nccl.json", 8)):
        line = _last_json_line(rel)
        assert line["n_gpus"] == n and line["scaling"] == "weak"
        assert line["config"]["reduction"] == "deferred by one replay"
        assert "sync_collective" in line and line["sync_collective"]["ms_per_step"] > line["ms_per_step"]
    eq = _last_json_line("profiles/r02/multi_gpu_equivalence_n8_all_modes.json")
    assert eq["ok"] and eq["world"] == 8
    assert max(v for k, v in eq.items() if k.startswith(("eager", "graph", "nvls", "deferred")) and isinstance(v, float)) < 2e-5
