"""tf.train: nothing the pinned path calls (the dump script differentiates with torch and saves with this repository's
own checkpoint writer)"""
