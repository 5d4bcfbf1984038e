"""Extracts the known-answer table of the reference's characteristic-function driver (test_crs_chf.f90:232-271,
`get_reference_val`, 32 complex values E[exp(i*omega_k*mean_j exp(X_j))], omega_k = k*pi/300) into chf_table.json.
The literals pass through default-kind `cmplx()` there, i.e. they are rounded to single precision: the JSON keeps the
printed literals and the tests round them the same way. Run in the build container only (needs /root/reference)."""
import json, re, pathlib

src = pathlib.Path("/root/reference/test_crs_chf.f90").read_text()
rows = re.findall(r"case\((\d+)\); val = cmplx\(([-+0-9.eE]+), ([-+0-9.eE]+)\)", src)
assert [int(k) for k, _, _ in rows] == list(range(32))
out = {"source": "test_crs_chf.f90:238-269", "omega_step": "pi/300", "values": [[float(a), float(b)] for _, a, b in rows]}
pathlib.Path(__file__).with_name("chf_table.json").write_text(json.dumps(out, indent=1))
