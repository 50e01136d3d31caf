"""Generate tests/golden/jpeg_golden.json in the build container (needs /root/reference):
  * Pillow's own files (image.save(format='JPEG', quality=q, optimize=o)) for small seeded images -- pins the
    codec behaviour to the library version the reference runs on here (Pillow / libjpeg-turbo are recorded);
  * the bytes the REFERENCE's ImagePreprocessor.compress_for_azure returns for a few images and targets that
    exercise the quality ladder and the resize fall-back.
Files are stored as sha256 + size (the tests regenerate the inputs from tests/golden/jpeg_images.py)."""
import hashlib, io, json, os, sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/backend")
import PIL  # noqa: E402
from PIL import Image, features  # noqa: E402
from utils.image_preprocessing import ImagePreprocessor  # noqa: E402  (the reference itself)
from jpeg_images import KINDS, SHAPES, image  # noqa: E402

files = []
for (h, w) in SHAPES:
    for kind in KINDS:
        a = image(kind, h, w)
        for q in (95, 85, 50, 30, 5, 100):
            for opt in (False, True):
                b = io.BytesIO()
                Image.fromarray(a).save(b, format="JPEG", quality=q, optimize=opt)
                d = b.getvalue()
                files.append(dict(kind=kind, h=h, w=w, quality=q, optimize=opt, size=len(d), sha=hashlib.sha256(d).hexdigest()))
ip = ImagePreprocessor()
azure = []
for kind, h, w, target in [("noise", 93, 127, 0.02), ("noise", 93, 127, 0.008), ("noise", 121, 33, 0.002), ("photo", 93, 127, 0.004),
                           ("text", 50, 70, 2.0), ("smooth", 93, 127, 0.0015), ("noise", 300, 420, 0.05)]:
    d = ip.compress_for_azure(Image.fromarray(image(kind, h, w)), target_size_mb=target)
    azure.append(dict(kind=kind, h=h, w=w, target_size_mb=target, size=len(d), sha=hashlib.sha256(d).hexdigest()))
out = os.path.join(HERE, "jpeg_golden.json")
json.dump(dict(generator="tests/golden/make_jpeg_golden.py", pillow=PIL.__version__, libjpeg=features.version("jpg"),
               libjpeg_turbo=features.check_feature("libjpeg_turbo"), files=files, compress_for_azure=azure), open(out, "w"))
print(out, len(files), "files", len(azure), "azure cases", os.path.getsize(out), "bytes")
