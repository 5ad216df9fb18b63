"""Counter wire format: against the reference's own update_payload_with_state when the reference tree is
present (build container), and self-consistency (payload -> log -> restore) everywhere."""
import json

import numpy as np
import pytest

from deepdish_b200 import wire

LABELS = ["person", "bicycle", "car"]
COUNTS = np.array([[5, 2, 7, 1], [0, 3, 3, 0], [9, 9, 18, 4]], dtype=np.int64)


def test_payload_log_restore_roundtrip(tmp_path):
    p = wire.counts_payload(COUNTS, LABELS)
    assert p["poscount_person"] == 5 and p["negcount_person"] == 2 and p["diff_person"] == 3
    assert p["intcount_car"] == 18 and p["delcount_car"] == 4 and p["diff_bicycle"] == -3
    assert list(p)[:5] == ["poscount_person", "negcount_person", "diff_person", "intcount_person", "delcount_person"]
    log = tmp_path / "counts.log"
    log.write_text(wire.log_line(COUNTS * 0, LABELS, 1000.0, 1) + wire.log_line(COUNTS, LABELS, 1001.0, 42))
    counts, frames = wire.restore_from_log(str(log), LABELS)
    assert counts == COUNTS.tolist() and frames == 42
    empty = tmp_path / "empty.log"
    empty.write_text("")
    assert wire.restore_from_log(str(empty), LABELS) == ([[0, 0, 0, 0]] * 3, 0)
    ev = json.loads(wire.crossing_event(COUNTS, LABELS, 12.5, "cam1", "pos"))
    assert ev["acp_event"] == "crossing" and ev["acp_event_value"] == "pos" and ev["acp_ts"] == "12.5"
    assert json.loads(wire.heartbeat_event(COUNTS, LABELS, 3.0, "cam1"))["acp_event"] == "heartbeat"


def test_payload_matches_reference_pipeline():
    from oracle import refload
    if not refload.available():
        pytest.skip("reference tree not present")
    mod = refload.load_pipeline_module()
    p = object.__new__(mod.Pipeline)
    p.wanted_labels = LABELS
    p.poscount = {l: int(COUNTS[i, 0]) for i, l in enumerate(LABELS)}
    p.negcount = {l: int(COUNTS[i, 1]) for i, l in enumerate(LABELS)}
    p.intcount = {l: int(COUNTS[i, 2]) for i, l in enumerate(LABELS)}
    p.delcount = {l: int(COUNTS[i, 3]) for i, l in enumerate(LABELS)}
    ref = {}
    p.update_payload_with_state(ref)
    got = wire.counts_payload(COUNTS, LABELS)
    assert got == ref and list(got) == list(ref)
