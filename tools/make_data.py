#!/usr/bin/env python
"""Derive the compact numeric tables the hot path's feeder needs from the
reference's data files (run once in the build container; the outputs are
committed because /root/reference does not exist on the GPU box).

Inputs (read-only, never edited):
  /root/reference/gadfly/data/hyperparameters.json              (5 granulation + 4 per-degree p-mode fits)
  /root/reference/gadfly/data/broomhall2009_table2_labeled.ecsv (81 BiSON p-mode frequencies + degree)
  /root/reference/notebooks/huber2011.ecsv                      (Huber et al. 2011 Kepler star table)

Outputs:
  gadfly_b200/data/solar_fit.json      {"granulation": [[S0,w0,Q]...], "p_mode_S0": [l0..l3],
                                        "p_mode_Q": [l0..l3], "bison_nu_uHz": [...], "bison_degree": [...]}
  gadfly_b200/data/huber2011_stars.csv KIC,mass,sig_mass,rad,sig_rad,teff,sig_teff,lum,sig_lum,numax,delta_nu
"""
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(HERE, "gadfly_b200", "data")


def read_ecsv(path):
    """Minimal ECSV reader: skip '#' header lines, first data line = column names."""
    rows, names = [], None
    with open(path) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            if names is None:
                names = line.split()
                continue
            rows.append(line.split())
    return names, rows


def main():
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(REF, "gadfly", "data", "hyperparameters.json")) as fh:
        hp = json.load(fh)
    gran = [[p["hyperparameters"][k] for k in ("S0", "w0", "Q")]
            for p in hp if p["metadata"]["source"] == "granulation"]
    osc = sorted((p for p in hp if p["metadata"]["source"] == "oscillation"),
                 key=lambda p: p["metadata"]["degree"])
    names, rows = read_ecsv(os.path.join(REF, "gadfly", "data", "broomhall2009_table2_labeled.ecsv"))
    assert names == ["nu", "degree"]
    solar = {
        "source": "derived from bmorris3/gadfly data/hyperparameters.json and "
                  "data/broomhall2009_table2_labeled.ecsv (Broomhall et al. 2009, Table 2)",
        "units": {"S0": "ppm^2/uHz", "w0": "rad*uHz", "nu": "uHz"},
        "granulation": gran,
        "p_mode_S0": [p["hyperparameters"]["S0"] for p in osc],
        "p_mode_Q": [p["hyperparameters"]["Q"] for p in osc],
        "bison_nu_uHz": [float(r[0]) for r in rows],
        "bison_degree": [int(r[1]) for r in rows],
    }
    with open(os.path.join(OUT, "solar_fit.json"), "w") as fh:
        json.dump(solar, fh, indent=1)

    names, rows = read_ecsv(os.path.join(REF, "notebooks", "huber2011.ecsv"))
    keep = ["KIC", "mass", "sig_mass", "rad", "sig_rad", "teff", "sig_teff", "lum", "sig_lum",
            "numax", "delta_nu"]
    idx = [names.index(k) for k in keep]
    with open(os.path.join(OUT, "huber2011_stars.csv"), "w") as fh:
        fh.write("# Huber et al. (2011, ApJ 743, 143) Kepler stars; columns derived from "
                 "bmorris3/gadfly notebooks/huber2011.ecsv\n")
        fh.write(",".join(keep) + "\n")
        for r in rows:
            fh.write(",".join(r[i] for i in idx) + "\n")
    print("wrote", OUT, len(solar["bison_nu_uHz"]), "modes,", len(rows), "stars")


if __name__ == "__main__":
    main()
