"""The one real location table the reference ships (data/helsinki/locations.csv.gz: 69 508 locations) as a compact
fixture for parity tests on real data: tests/golden/helsinki_locations.npz with the columns the loader reads
(factory/ConstructHerosModel.java:300-395: location_category -> type id, nb_sublocations, area per sub-location, lat,
lon).  Run in the build container (the reference tree does not travel to the GPU box); the .npz is committed."""
import csv
import gzip
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from heros_b200.scenario import TYPE_ID  # noqa: E402

with gzip.open("/root/reference/data/helsinki/locations.csv.gz", "rt") as f:
    rows = list(csv.reader(f))
h = rows[0]
col = {name: h.index(name) for name in ("location_id", "nb_sublocations", "location_category", "area", "lat", "lon")}
body = rows[1:]
assert [int(r[col["location_id"]]) for r in body] == list(range(len(body)))  # ids are the row numbers
out = dict(
    type=np.array([TYPE_ID[r[col["location_category"]]] for r in body], np.uint8),
    nsub=np.array([int(float(r[col["nb_sublocations"]])) for r in body], np.int16),
    area_per_sub=np.array([float(r[col["area"]]) for r in body], np.float32),
    lat=np.array([float(r[col["lat"]]) for r in body], np.float32),
    lon=np.array([float(r[col["lon"]]) for r in body], np.float32))
path = os.path.join(ROOT, "tests", "golden", "helsinki_locations.npz")
np.savez_compressed(path, **out)
print(len(body), int(out["nsub"].sum()), int((out["nsub"] == 0).sum()), int((out["area_per_sub"] < 0).sum()),
      os.path.getsize(path))
