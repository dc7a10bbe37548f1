"""The host programs' embedding-file reader / writer (kb2e_b200/host/loader.cpp) is byte-compatible with the reference's
fprintf("%.6lf\\t") / fscanf("%lf") (common/trainer.cpp:109-127, common/evaluation.cpp:74-105): CPU only."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_driver(tmp_path):
    exe = str(tmp_path / "io_driver")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "kb2e_b200", "host"), "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "host", "io_driver.cpp"), os.path.join(ROOT, "kb2e_b200", "host", "loader.cpp"), "-o", exe],
                   check=True)
    return exe


def test_table_text_round_trip_is_byte_compatible(tmp_path):
    exe = build_driver(tmp_path)
    rng = np.random.default_rng(0)
    rows, cols = 257, 50
    x = rng.normal(0, 0.3, (rows, cols))
    # rounding ties at the 6th decimal, tiny, negative zero, large and huge magnitudes
    x[0, :8] = [0.0000005, -0.0000005, 0.1234565, 2.5e-7, -0.0, 123456789.123456789, -1e-300, 1e22]
    x[1, :4] = [1.0000015, 0.9999995, -0.9999995, 1e300]
    src, txt, out = tmp_path / "in.bin", tmp_path / "table.txt", tmp_path / "out.bin"
    x.astype(np.float64).tofile(src)
    subprocess.run([exe, str(rows), str(cols), str(src), str(txt), str(out)], check=True)
    want = "".join("".join("%.6f\t" % v for v in row) + "\n" for row in x)
    got = open(txt).read()
    assert got == want
    back = np.fromfile(out, dtype=np.float64).reshape(rows, cols)
    expect = np.array([[float(tok) for tok in line.split("\t")[:-1]] for line in want.splitlines()])
    assert np.array_equal(back, expect)
    assert np.signbit(back[0, 4])  # "-0.000000" keeps its sign like fscanf


def test_loader_accepts_what_fscanf_accepts(tmp_path):
    exe = build_driver(tmp_path)
    # write with the driver, then replace the text with spellings strtod / fscanf accept and parse again through a 1 x 6 table
    vals = np.array([[0.0] * 6])
    src, txt, out = tmp_path / "in.bin", tmp_path / "t.txt", tmp_path / "out.bin"
    vals.tofile(src)
    subprocess.run([exe, "1", "6", str(src), str(txt), str(out)], check=True)
    # loadTable is exercised directly by a second tiny run: the driver rewrites the file, so check the parser through a
    # copy of its logic instead: feed the odd spellings as the INPUT doubles' text form via a pre-written file
    odd = "+1.5\t1e-3\t  -2.25E+2\t0x1p-1\t.5\t7.\t\n"
    open(txt, "w").write(odd)
    helper = tmp_path / "parse.cpp"
    helper.write_text('#include <cstdio>\n#include <vector>\n#include "loader.h"\nint main(int c, char** v) { std::vector<double> t; '
                      'if (!kb2e_host::loadTable(v[1], 1, 6, t)) return 1; for (double x : t) printf("%.17g\\n", x); return 0; }\n')
    exe2 = str(tmp_path / "parse")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "kb2e_b200", "host"), "-I" + os.path.join(ROOT, "include"),
                    str(helper), os.path.join(ROOT, "kb2e_b200", "host", "loader.cpp"), "-o", exe2], check=True)
    p = subprocess.run([exe2, str(txt)], capture_output=True, text=True, check=True)
    got = [float(s) for s in p.stdout.split()]
    assert got == [1.5, 1e-3, -225.0, 0.5, 0.5, 7.0]
