"""placeholder: the reference imports skimage.measure (utils/eval_utils.py:3) but calls only skimage.metrics"""
