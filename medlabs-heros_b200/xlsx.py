"""A minimal reader of .xlsx workbooks (zip + SpreadsheetML), enough for the activity-schedule workbooks of the
reference (data/<city>/activities/*.xlsx) and its result spreadsheets; no third-party dependency."""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
import zipfile
from typing import Dict, List, Optional

MAIN = "http://schemas.openxmlformats.org/spreadsheetml/2006/main"
REL = "http://schemas.openxmlformats.org/officeDocument/2006/relationships"
NS = {"m": MAIN}


def _column_index(ref: str) -> int:
    index = 0
    for ch in re.match(r"[A-Z]+", ref).group(0):
        index = index * 26 + ord(ch) - 64
    return index - 1


def read_xlsx(path: str) -> Dict[str, List[List[Optional[str]]]]:
    """sheet name -> rows -> cell texts (None for empty cells); numbers come as the strings the file stores"""
    z = zipfile.ZipFile(path)
    shared: List[str] = []
    if "xl/sharedStrings.xml" in z.namelist():
        root = ET.fromstring(z.read("xl/sharedStrings.xml"))
        for si in root.findall("m:si", NS):
            shared.append("".join(t.text or "" for t in si.iter("{%s}t" % MAIN)))
    workbook = ET.fromstring(z.read("xl/workbook.xml"))
    rels = ET.fromstring(z.read("xl/_rels/workbook.xml.rels"))
    target = {r.get("Id"): r.get("Target") for r in rels}
    out: Dict[str, List[List[Optional[str]]]] = {}
    for sheet in workbook.find("m:sheets", NS):
        t = target[sheet.get("{%s}id" % REL)]
        t = t[1:] if t.startswith("/") else "xl/" + t
        rows = []
        for row in ET.fromstring(z.read(t)).iter("{%s}row" % MAIN):
            cells = {}
            for c in row.findall("m:c", NS):
                v = c.find("m:v", NS)
                if v is None:
                    inline = c.find("m:is", NS)
                    val = "".join(x.text or "" for x in inline.iter("{%s}t" % MAIN)) if inline is not None else None
                elif c.get("t") == "s":
                    val = shared[int(v.text)]
                else:
                    val = v.text
                cells[_column_index(c.get("r"))] = val
            rows.append([cells.get(i) for i in range(max(cells) + 1)] if cells else [])
        out[sheet.get("name")] = rows
    return out
