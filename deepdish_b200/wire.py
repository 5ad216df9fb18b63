"""Wire format of the counters: the reduced count tensor i64 [C,4] (pos, neg, int,
del per label) <-> the reference's MQTT / JSON-lines payload keys, and restore-from-log.

Host-side serialisation only -- follows deepdish.py:1141-1145 (update_payload_with_state), :1155-1166
(crossing event / log line), :1168-1183 (heartbeat) and :545-561 (--restore-from-log).
"""
import json
from collections import deque
from time import asctime, localtime


def counts_payload(counts, labels):
    """deepdish.py:1141-1145 -- poscount_/negcount_/diff_/intcount_/delcount_<label> in label order."""
    rows = counts.tolist() if hasattr(counts, "tolist") else counts
    out = {}
    for lbl, (pos, neg, inter, dele) in zip(labels, rows):
        out['poscount_' + lbl] = int(pos)
        out['negcount_' + lbl] = int(neg)
        out['diff_' + lbl] = int(pos) - int(neg)
        out['intcount_' + lbl] = int(inter)
        out['delcount_' + lbl] = int(dele)
    return out


def crossing_event(counts, labels, t_frame, acp_id, crossing_type, temp=None):
    """MQTT 'crossing' payload, deepdish.py:1155-1159."""
    payload = {'acp_ts': str(t_frame), 'acp_id': acp_id, 'acp_event': 'crossing', 'acp_event_value': crossing_type,
               'temp': temp}
    payload.update(counts_payload(counts, labels))
    return json.dumps(payload)


def heartbeat_event(counts, labels, now, acp_id, temp=None):
    """MQTT 'heartbeat' payload, deepdish.py:1171-1175."""
    payload = {'acp_ts': str(now), 'acp_id': acp_id, 'acp_event': 'heartbeat', 'temp': temp}
    payload.update(counts_payload(counts, labels))
    return json.dumps(payload)


def log_line(counts, labels, t_frame, frame_count, temp=None):
    """JSON-lines log entry, deepdish.py:1161-1166."""
    payload = {'timestamp': str(t_frame), 'asctime': asctime(localtime(t_frame)), 'frame_count': frame_count,
               'temp': temp}
    payload.update(counts_payload(counts, labels))
    return json.dumps(payload) + '\n'


def restore_from_log(path, labels):
    """deepdish.py:545-561 -- counters [C][4] (pos, neg, int, del) and frame_count from the LAST log line;
    missing keys default to 0; an empty file gives zeros."""
    counts = [[0, 0, 0, 0] for _ in labels]
    frame_count = 0
    with open(path, mode='r') as f:
        q = deque(f, 1)
    if len(q) > 0:
        data = json.loads(q.pop())
        for i, lbl in enumerate(labels):
            counts[i] = [data.get('poscount_' + lbl, 0), data.get('negcount_' + lbl, 0),
                         data.get('intcount_' + lbl, 0), data.get('delcount_' + lbl, 0)]
        frame_count = data.get('frame_count', 0)
    return counts, frame_count
