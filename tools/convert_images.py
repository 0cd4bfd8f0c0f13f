"""Convert the images an NVM file names (JPEG/PNG/...) to the PPM files `tmvs` reads: for `dir/img0001.jpg` it writes
`dir/img0001.ppm` next to it (tmvs falls back to <stem>.ppm / <stem>.pgm when the named file is not PNM).
usage: python tools/convert_images.py scene.nvm"""
import os
import sys

import cv2


def main(nvm):
    base = os.path.dirname(os.path.abspath(nvm))
    lines = [l.split() for l in open(nvm) if l.strip()]
    assert lines[0][0] == "NVM_V3", "not an NVM_V3 file"
    n = int(lines[1][0])
    for l in lines[2:2 + n]:
        src = os.path.join(base, l[0])
        img = cv2.imread(src, cv2.IMREAD_COLOR)
        if img is None:
            print("cannot read", src)
            continue
        dst = os.path.splitext(src)[0] + ".ppm"
        cv2.imwrite(dst, img)
        print(dst)


if __name__ == "__main__":
    main(sys.argv[1])
