"""Regenerates tests/golden/ from the reference checkout (run in the build container only; the GPU box has no
/root/reference).  Copies the reference's golden ARCHIVES (test data, not source) for the hot path and writes
manifest.json with, per archive, the derived keys (so tests need no Argon2 on the hot path) and the SHA-256 of
every file entry as decoded by the CPU oracle -- which this script first checks against resources/test/raw.
"""
import hashlib, json, os, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pna_oracle as O  # noqa: E402

REF = "/root/reference/resources/test"
FIXTURES = ["deflate.pna", "zstd.pna", "zstd_aes_ctr.pna", "zstd_aes_cbc.pna", "zstd_camellia_ctr.pna",
            "zstd_camellia_cbc.pna", "zstd_with_raw_file_size.pna", "zstd_keep_xattr.pna", "zstd_keep_all.pna",
            "zstd_keep_dir.pna", "zstd_keep_permission.pna", "zstd_keep_timestamp.pna", "solid_zstd.pna",
            "solid_deflate.pna", "solid_zstd_aes_ctr.pna", "solid_zstd_aes_cbc.pna", "solid_zstd_camellia_ctr.pna",
            "solid_zstd_camellia_cbc.pna", "solid_zstd_keep_all.pna", "empty.pna", "multipart.part1.pna",
            "multipart.part2.pna", "zstd_aes_gcm.pna", "zstd_camellia_gcm.pna", "solid_zstd_aes_gcm.pna",
            "solid_zstd_camellia_gcm.pna", "xz.pna", "solid_xz.pna", "0.33.0/zstd_keep_all.pna", "zstd_keep_fflags.pna"]
PASSWORD = b"password"


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    raw = {}
    for root, _, files in os.walk(os.path.join(REF, "raw")):
        for f in files:
            p = os.path.join(root, f)
            raw[os.path.relpath(p, REF)] = open(p, "rb").read()
    manifest = {"password": PASSWORD.decode(), "raw_sha256": {k: hashlib.sha256(v).hexdigest() for k, v in raw.items()},
                "archives": {}}
    # icon.bmp is a missing large blob in the checkout; three fixtures agree on its hash
    for fx in FIXTURES:
        src = os.path.join(REF, fx)
        dst = os.path.join(here, "ref", fx.replace("/", "__"))
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        buf = open(src, "rb").read()
        info = {"file": os.path.relpath(dst, here), "size": len(buf), "keys": {}, "entries": [], "expect": "ok"}
        keys = {}
        try:
            for name, data in O.extract_all(buf, PASSWORD, keys):
                if name in raw:
                    assert raw[name] == data, (fx, name)
                info["entries"].append({"name": name, "size": len(data), "sha256": hashlib.sha256(data).hexdigest()})
        except O.OracleError as e:
            info["expect"] = {O.UNSUPPORTED: "unsupported"}.get(e.status, f"error{e.status}")
        info["keys"] = {k: v.hex() for k, v in keys.items()}
        manifest["archives"][fx] = info
        print(fx, info["expect"], len(info["entries"]))
    shutil.copyfile(os.path.join(REF, "multipart_test.txt"), os.path.join(here, "ref", "multipart_test.txt"))
    json.dump(manifest, open(os.path.join(here, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
