"""Reference arm of bench.py: run the UNMODIFIED reference (Microsoft/multiverso) installed under
``baseline/_ref`` through its own CLI (Applications/WordEmbedding) on the same metric/config.

The reference is a CMake/MPI C++ project: ``pip install /root/reference`` cannot work (there
is no setup.py / pyproject at its root) and CMake needs ``find_package(MPI REQUIRED)``
(CMakeLists.txt:11) while the image has no MPI.  See DESIGN.md "Reference arm" for the recorded
outcome.  ``tools/build_reference.sh`` attempts an out-of-tree build of the unmodified sources
against a tiny single-node MPI shim (the reference only touches 15 MPI symbols); when that
build exists at baseline/_ref/bin/wordembedding this arm runs it.
"""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "baseline", "_ref", "bin", "wordembedding")


def unavailable(why: str) -> dict:
    return {"impl": "reference", "unavailable": why}


def run(args) -> dict:
    if not os.path.exists(REF_BIN) and os.path.isdir("/root/reference/src"):
        # build on the fly where the reference sources are mounted (seconds; unmodified sources)
        try:
            subprocess.run(["bash", os.path.join(ROOT, "tools", "build_reference.sh")], capture_output=True,
                           timeout=600, check=False)
        except Exception:
            pass
    if not os.path.exists(REF_BIN):
        return unavailable("reference needs MPI/ZeroMQ (CMakeLists.txt:11 find_package(MPI REQUIRED)); "
                           "neither exists in this image and pip cannot install a CMake C++ project "
                           "without setup.py; no baseline/_ref/bin/wordembedding was built")
    bw_bin = os.path.join(os.path.dirname(REF_BIN), "matrix_bw")
    try:
        from baseline import reference_runner
        if getattr(args, "metric", "words") == "matrix_bw":
            if not os.path.exists(bw_bin):
                return unavailable("baseline/_ref/bin/matrix_bw was not built (tools/build_reference.sh)")
            out = reference_runner.run_matrix_bw(bw_bin)
            if "metric" in out:
                world = int(os.environ.get("WORLD_SIZE", "1"))
                out.update({"n_gpus": world, "steps": out["iters"], "warmup": 1, "ms_per_step": out["add_ms"] + out["get_ms"],
                            "scaling": "strong", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                            "impl": "reference", "gpu_launches": 0,
                            "e2e": {"value": out["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                    "note": "CPU program: host buffers in, host buffers out -- end to end by construction"}})
            return out
        out = reference_runner.run_wordembedding(REF_BIN, args)
        # second half of the BASELINE.json metric: MatrixTable Get+Add GB/s through the same unmodified library
        if not getattr(args, "no_table_bw", False) and os.path.exists(bw_bin):
            try:
                bw = reference_runner.run_matrix_bw(bw_bin)
                if "metric" in out and "metric" in bw:
                    out["secondary"] = bw
            except Exception as e:  # the headline metric must still print
                if "metric" in out:
                    out["secondary"] = {"metric": "matrix_table_get_plus_add_gbs", "unavailable": repr(e)[:200]}
        return out
    except Exception as e:  # never crash the driver
        return unavailable(f"reference run failed: {e!r}"[:300])
