"""Generate tests/golden/resnet18_kat.json: semantic known-answer test of the ORACLE's ResNet-18.

Runs in the build container only (needs /root/reference for the reference's own resnet18.npy, test
images and imagenet-classes.txt; reference resnet.py:453-488 is the smoke test this mirrors).
Preprocessing follows resnet.py:111-121 + preprocessing.py:135-172: central crop 0.875, bilinear resize
to 224x224 (align_corners=False), x/255, ImageNet mean/std.  BN runs in moving-statistics mode here
(is_training=False) -- this validates HWIO layout, TF-SAME padding, max-pool and block wiring of the
oracle; the batch-statistics branch used by the hot path is covered by the analytic tests.
"""
import json, os, sys
import numpy as np
import torch
import torch.nn.functional as F
from PIL import Image

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from oracle import sag_oracle as O

REF = '/root/reference/pyutils/tflib/models/image'


def central_crop(img, frac):                      # tf.image.central_crop semantics
    h, w = img.shape[:2]
    y0 = int((h - h * frac) / 2)
    x0 = int((w - w * frac) / 2)
    return img[y0:h - y0, x0:w - x0]


def resize_bilinear_tf(img, oh, ow):              # tf.image.resize_bilinear(align_corners=False), legacy (no half-pixel)
    t = torch.as_tensor(img).permute(2, 0, 1)[None].float()
    h, w = t.shape[2:]
    ys = torch.arange(oh) * (h / oh)
    xs = torch.arange(ow) * (w / ow)
    y0 = ys.floor().long().clamp(max=h - 1); y1 = (y0 + 1).clamp(max=h - 1); fy = (ys - y0).view(1, 1, -1, 1)
    x0 = xs.floor().long().clamp(max=w - 1); x1 = (x0 + 1).clamp(max=w - 1); fx = (xs - x0).view(1, 1, 1, -1)
    top = t[:, :, y0][:, :, :, x0] * (1 - fx) + t[:, :, y0][:, :, :, x1] * fx
    bot = t[:, :, y1][:, :, :, x0] * (1 - fx) + t[:, :, y1][:, :, :, x1] * fx
    return (top * (1 - fy) + bot * fy)[0].permute(1, 2, 0)


def main():
    pre = np.load(os.path.join(REF, 'resnet18.npy'), allow_pickle=True, encoding='latin1').item()
    W = O.Weights(pre)
    classes = [l.strip() for l in open(os.path.join(REF, 'imagenet-classes.txt'))]
    out = {}
    for fn in sorted(os.listdir(os.path.join(REF, 'test_images'))):
        if not fn.lower().endswith(('.jpg', '.jpeg', '.png')):
            continue
        img = np.asarray(Image.open(os.path.join(REF, 'test_images', fn)).convert('RGB'))
        x = resize_bilinear_tf(central_crop(img, 0.875), 224, 224) / 255.
        x = (x - torch.tensor([0.485, 0.456, 0.406])) / torch.tensor([0.229, 0.224, 0.225])
        logits, _ = O.resnet18(W, '', x[None], bn_train=False, truncate_at=None)
        top = torch.argsort(-logits[0])[:5].tolist()
        out[fn] = {'top5': top, 'top5_names': [classes[i] for i in top],
                   'top1_logit': float(logits[0, top[0]])}
        print(fn, out[fn]['top5_names'])
    with open(os.path.join(os.path.dirname(__file__), 'resnet18_kat.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
