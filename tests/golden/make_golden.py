"""Writes tests/golden/golden_4x4.json: the reference's only known-answer vector, transcribed from
test/src/test_shared_loop.cpp:15-34 (input rows [1,2,3,4,0,0], expected row 0 [40,0,-8,8,-8,0]).
The reference itself cannot be built or imported here (C++/HPX/FFTW), so the vector is transcribed,
not generated; oracle/oracle.py reproduces it exactly (tests/test_oracle.py)."""
import json
import os

inp = [[1.0, 2.0, 3.0, 4.0, 0.0, 0.0] for _ in range(4)]
exp = [[40.0, 0.0, -8.0, 8.0, -8.0, 0.0]] + [[0.0] * 6 for _ in range(3)]
with open(os.path.join(os.path.dirname(__file__), "golden_4x4.json"), "w") as f:
    json.dump({"source": "test/src/test_shared_loop.cpp:15-34,53", "n_row": 4, "n_col": 6, "input": inp, "expected": exp}, f, indent=1)
