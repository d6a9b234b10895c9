"""Extract the two test patterns BASELINE.md section 4's anchors were recorded on from the reference's own archive
(extra/test_output_images.zip, original/601cb.png and original/cbar.png -- 562 and 2099 bytes) into tests/golden/inputs/,
so that the anchor tests also run where /root/reference is not mounted (the GPU box).  Run here, commit the result."""
import hashlib
import os
import zipfile

ZIP = "/root/reference/extra/test_output_images.zip"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    z = zipfile.ZipFile(ZIP)
    for name in ("601cb.png", "cbar.png"):
        data = z.read("test_output_images/original/" + name)
        with open(os.path.join(HERE, "inputs", name), "wb") as f:
            f.write(data)
        print(name, len(data), hashlib.sha256(data).hexdigest())
