"""The rBRIEF 31x31 pattern is a constant of the algorithm; pin it."""
import hashlib
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
SHA = "2164181aea6ff9ac426ca512d5130d15e1f6e3cd47b1cbdd568bbe1e55d49023"


def _table():
    txt = open(os.path.join(HERE, "..", "include", "hyorb_brief_pattern.inc")).read().split("*/", 1)[1]
    return [int(t) for t in txt.replace("\n", "").split(",") if t.strip()]


def test_pattern_checksum():
    v = _table()
    assert len(v) == 1024 and min(v) == -13 and max(v) == 12
    assert hashlib.sha256(bytes([(n + 256) % 256 for n in v])).hexdigest() == SHA


def test_pattern_equals_reference_table_when_available():
    ref = "/root/reference/src/features/low_level/ORBFinder.cpp"
    if not os.path.exists(ref):       # the GPU box has no /root/reference
        return
    body = "\n".join(open(ref).read().split("\n")[150:409])
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).split("=", 1)[1]
    assert [int(x) for x in re.findall(r"-?\d+", body)] == _table()
