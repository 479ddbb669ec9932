"""Golden vectors of the Otsu binarisation branch (DT_BIN_METHOD = "otsu", Detect_OBB.py:109-111), written by the
REFERENCE ITSELF (lifted functions, cv2 4.13 IPP off, numpy 2.3) in the build container:

  python tests/golden/make_otsu_golden.py  ->  tests/golden/otsu_golden.npz

out_<key>: build_multich(in_<key> of pixel_golden.npz, 4) with DT_BIN_METHOD = "otsu"; thr_<key>: the Otsu threshold
cv2 chose for that crop (the reference discards it; kept so a mismatch can be localised).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cv2  # noqa: E402

cv2.ipp.setUseIPP(False)
from oracle import lift_reference as LR  # noqa: E402


def main():
    ref = LR.load_detect(4)
    ref.DT_BIN_METHOD = "otsu"
    src = np.load(os.path.join(HERE, "pixel_golden.npz"))
    out = {}
    for name in src.files:
        if not name.startswith("in_"):
            continue
        key = name[3:]
        crop = np.ascontiguousarray(src[name])
        out["out_" + key] = ref.build_multich(crop, 4)
        # the threshold, through the same library calls the reference makes
        gray = cv2.cvtColor(crop, cv2.COLOR_BGR2GRAY)
        acc = None
        for s in ref.MS_SIGMAS:
            blur = cv2.GaussianBlur(gray, (0, 0), s, s, borderType=cv2.BORDER_REFLECT_101) if s > 0 else gray
            mag = cv2.magnitude(cv2.Scharr(blur, cv2.CV_32F, 1, 0), cv2.Scharr(blur, cv2.CV_32F, 0, 1))
            acc = mag if acc is None else np.maximum(acc, mag)
        acc8 = cv2.normalize(acc, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
        out["thr_" + key] = np.int32(cv2.threshold(acc8, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[0])
    np.savez_compressed(os.path.join(HERE, "otsu_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
