"""Writes tests/golden/params_alpha_area.properties and params_alpha_distance.properties (Java .properties syntax) from the committed
parameter fixtures tests/golden/params_alpha_{area,distance}.json (tools/make_golden_params.py made those from
src/main/resources/alpha-area.properties / alpha-distance.properties of the reference): the files the C++ host layer
(include/heros_host.hpp, tests/cpp/test_host.cpp) loads with its own .properties reader."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

if __name__ == "__main__":
    for model in ("area", "distance"):
        with open(os.path.join(GOLDEN, f"params_alpha_{model}.json")) as f:
            props = json.load(f)["properties"]
        path = os.path.join(GOLDEN, f"params_alpha_{model}.properties")
        with open(path, "w") as f:
            f.write(f"# parameter values of the reference's alpha-{model}.properties (see tools/make_properties_fixtures.py)\n")
            f.write(f"generic.diseasePropertiesModel = {model}\n! the seed of the shipped scenarios\ngeneric.Seed = 111\n")
            for k in sorted(props):
                f.write(f"{k} = {props[k]}\n")
        print(path, len(props), "keys")
