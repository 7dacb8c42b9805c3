"""imageio entry points the reference's utils/vis_utils.py uses (mimsave / imwrite / imread), backed by Pillow."""
import numpy as np
from PIL import Image


def mimsave(uri, ims, format=None, duration=0.1, **kwargs):
    frames = [Image.fromarray(np.asarray(im)) for im in ims]
    if not frames:
        return
    ms = int(round(1000 * duration)) if duration < 10 else int(duration)      # imageio v2: seconds, v3: milliseconds
    frames[0].save(uri, save_all=True, append_images=frames[1:], duration=ms, loop=0)


def imwrite(uri, im, **kwargs):
    Image.fromarray(np.asarray(im)).save(uri)


imsave = imwrite


def imread(uri, **kwargs):
    return np.asarray(Image.open(uri))
