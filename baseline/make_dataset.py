"""A synthetic on-disk dataset in the layout datasets/voc.py:36-39,95 reads:

    <root>/JPEGImages/<name>.jpg               RGB, ~500x375 like VOC
    <root>/SegmentationClassAug/<name>.png     uint8 class indices (0 = background, 255 = border)
    <root>/lists/{train_aug,val}.txt           image names
    <root>/lists/cls_labels_onehot.npy         {name: float32[20]} (pickled dict, as the reference's own file)

Images are smoothed noise with 1-3 coloured ellipses, one per "object class"; the label PNG marks them.  Seeded.
"""
import os

import numpy as np
from PIL import Image


def make_voc_like(root, n_train=32, n_val=4, seed=0, size=(375, 500), num_classes=21):
    rng = np.random.RandomState(seed)
    img_dir, lab_dir, list_dir = (os.path.join(root, d) for d in ("JPEGImages", "SegmentationClassAug", "lists"))
    for d in (img_dir, lab_dir, list_dir):
        os.makedirs(d, exist_ok=True)
    H, W = size
    yy, xx = np.mgrid[0:H, 0:W]
    names = {"train_aug": [], "val": []}
    onehot = {}
    for i in range(n_train + n_val):
        split = "train_aug" if i < n_train else "val"
        name = f"2099_{i:06d}"
        base = rng.randint(0, 256, (H // 8 + 1, W // 8 + 1, 3)).astype(np.uint8)
        img = np.asarray(Image.fromarray(base).resize((W, H), resample=Image.BILINEAR)).astype(np.float32)
        label = np.zeros((H, W), np.uint8)
        vec = np.zeros(num_classes - 1, np.float32)
        for k in rng.choice(num_classes - 1, rng.randint(1, 4), replace=False):
            cy, cx = rng.randint(H // 4, 3 * H // 4), rng.randint(W // 4, 3 * W // 4)
            ry, rx = rng.randint(H // 8, H // 3), rng.randint(W // 8, W // 3)
            inside = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
            colour = np.array([(37 * (k + 1)) % 256, (91 * (k + 3)) % 256, (153 * (k + 7)) % 256], np.float32)
            img[inside] = 0.6 * colour + 0.4 * img[inside]
            label[inside] = k + 1
            vec[k] = 1.0
        Image.fromarray(img.clip(0, 255).astype(np.uint8)).save(os.path.join(img_dir, name + ".jpg"), quality=90)
        Image.fromarray(label).save(os.path.join(lab_dir, name + ".png"))
        names[split].append(name)
        onehot[name] = vec
    for split, lst in names.items():
        with open(os.path.join(list_dir, split + ".txt"), "w") as f:
            f.write("\n".join(lst) + "\n")
    np.save(os.path.join(list_dir, "cls_labels_onehot.npy"), onehot, allow_pickle=True)
    return root, list_dir


if __name__ == "__main__":
    import sys
    print(make_voc_like(sys.argv[1] if len(sys.argv) > 1 else "/tmp/dupl_voc_synth"))
