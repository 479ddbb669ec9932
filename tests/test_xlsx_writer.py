"""f4: the stdlib .xlsx writer used when pandas has no Excel engine (openpyxl is not installed here): the file
is a valid one-sheet workbook whose cells read back exactly, with the reference's columns (Detect_OBB.py:328)
and the same cell encoding as the reference's Output/*.xlsx (inline strings, plain numbers)."""
import json
import os
import xml.etree.ElementTree as ET
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
NS = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main"}


def _read(path):
    z = zipfile.ZipFile(path)
    assert {"[Content_Types].xml", "_rels/.rels", "xl/workbook.xml", "xl/_rels/workbook.xml.rels",
            "xl/worksheets/sheet1.xml"} <= set(z.namelist())
    rows = []
    for row in ET.fromstring(z.read("xl/worksheets/sheet1.xml")).find("m:sheetData", NS):
        vals = []
        for c in row:
            vals.append(c.find("m:is/m:t", NS).text if c.get("t") == "inlineStr" else float(c.find("m:v", NS).text))
        rows.append(vals)
    return rows


def test_xlsx_round_trip_of_reference_rows(tmp_path):
    import importlib.util
    # detect.py imports the CUDA library at import time; the writer itself is plain Python
    import __graft_entry__ as g
    g.build()
    from oriented_object_detection_b200 import detect
    with open(os.path.join(HERE, "golden", "xlsx_rows.json")) as fh:
        ref = json.load(fh)
    for name in ("Test1", "Test2"):
        cols, rows = ref[name]["columns"], ref[name]["rows"]
        assert cols == detect.XLSX_COLUMNS
        p = str(tmp_path / (name + ".xlsx"))
        detect._write_xlsx(p, cols, rows)
        back = _read(p)
        assert back[0] == cols
        assert back[1:] == [[r[0]] + [float(v) for v in r[1:]] for r in rows]      # repr() round-trips every float
