"""Input pipeline of nabu on in-process iterators.

reference: processing/input_pipeline.py -- `get_filenames` (:10-55), `input_pipeline` (:57-174: one reader per data
stream, `bucket_by_sequence_length` on the FIRST stream's length with `dynamic_pad`, optional variable batch size),
`bucket_boundaries` (:176-202).  TF's queue runners are replaced by a deterministic generator: examples are read in
the order of the (optionally shuffled) filename list, appended to their bucket, and a bucket is emitted when it holds
its batch size; leftovers are emitted at the end when `allow_smaller_final_batch` (the reference's queues emit in a
thread-dependent order, so only the batch CONTENTS per bucket are comparable, not the order of batches).
"""
import os

import numpy as np

from . import tfreaders


def get_filenames(dataconfs):
    """dataconfs: one list of database sections (dicts with 'dir') per data stream.  Returns (tab-joined file names
    per example, example names).  Keys are `<utterance>-<index of the section within its stream>`; an example is kept
    only when every stream has it, in the order of the first stream's pointers.scp (reference :10-55)."""
    def stream_table(sections):
        table = {}
        for idx, section in enumerate(sections):
            with open(os.path.join(section['dir'], 'pointers.scp')) as scp:
                for line in scp:
                    utt, path = line.strip().split('\t')
                    table['%s-%d' % (utt, idx)] = path
        return table

    tables = [stream_table(sections) for sections in dataconfs]
    elements, names = [], []
    for key, first in tables[0].items():
        missing = [t for t in tables[1:] if key not in t]
        if missing:
            print('%s was not found in all sets of data, ignoring this example' % key)
            continue
        elements.append('\t'.join([first] + [t[key] for t in tables[1:]]))
        names.append(key)
    return elements, names


def bucket_boundaries(histogram, numbuckets):
    """Greedy bucket boundaries that spread the utterances evenly (reference :176-202, reproduced decision for
    decision -- tests/golden/bucket_boundaries.json comes from the reference's own function): bucket i starts at the
    previous boundary and grows while one more length bin does not move its population further from the target
    (remaining utterances // remaining buckets).  Prefix sums replace the reference's repeated slice sums; the counts
    are integers, so the sums are exact either way."""
    counts = np.asarray(histogram, dtype=np.float64)
    nbins = counts.shape[0]
    prefix = np.concatenate([[0.0], np.cumsum(counts)])         # prefix[j] = counts[:j].sum()
    out, start = [], 0
    for made in range(numbuckets - 1):
        # with more buckets than the data can fill, `start` runs past the histogram one bin per bucket (the
        # reference's slices are empty there); its population is then zero
        base = prefix[min(start, nbins)]
        target = int((prefix[nbins] - base) / (numbuckets - made))
        if target == 0:
            print('%d buckets could not be reached, using %d buckets' % (numbuckets, made))
        edge = start + 1
        while edge + 1 < nbins and abs(prefix[edge] - base - target) >= abs(prefix[edge + 1] - base - target):
            edge += 1
        out.append(edge)
        start = edge
    return out


def batch_plan(histogram, batch_size, numbuckets, variable_batch_size=False):
    """(boundaries, batch size per bucket, num_steps) exactly as input_pipeline.py:131-160 computes them"""
    histogram = np.asarray(histogram)
    if numbuckets > 1:
        boundaries = bucket_boundaries(histogram, numbuckets)
        if variable_batch_size:
            batch_sizes = [max(int(batch_size * boundaries[0] / b), 1) for b in boundaries + [histogram.size]]
            numutt = [histogram[boundaries[i]:b].sum() for i, b in enumerate(boundaries[1:])]
            numutt = [histogram[:boundaries[0]].sum()] + numutt + [histogram[boundaries[-1]:].sum()]
            num_steps = int((np.array(numutt) / np.array(batch_sizes)).sum())
        else:
            batch_sizes = [int(batch_size)] * (len(boundaries) + 1)
            num_steps = int(histogram.sum() / int(batch_size))
        return boundaries, batch_sizes, num_steps
    return [], [int(batch_size)], int(histogram.sum() / int(batch_size))


def _pad(arrays):
    """tf's dynamic_pad: zero-pad every example to the longest of the batch along axis 0"""
    T = max(a.shape[0] for a in arrays)
    out = np.zeros((len(arrays), T) + arrays[0].shape[1:], arrays[0].dtype)
    for i, a in enumerate(arrays):
        out[i, :a.shape[0]] = a
    return out


class BatchSource(object):
    """Iterable of `(inputs, input_seq_length, targets, target_seq_length)` dict tuples of torch tensors -- what
    Trainer.train / Evaluator.evaluate consume.  `input_names` / `target_names` name the streams in the order of
    `dataconfs` (inputs first), as trainers/trainer.py:404-415 does."""

    def __init__(self, dataconfs, input_names, target_names, batch_size, numbuckets=1, variable_batch_size=False,
                 allow_smaller_final_batch=False, shuffle_seed=None, device='cpu', rank=0, world=1, prefetch=None):
        import torch
        self._torch = torch
        self.device = device
        self.input_names, self.target_names = list(input_names), list(target_names)
        self.elements, self.names = get_filenames(dataconfs)
        self.readers = []
        for dataconfset in dataconfs:
            types = [d['type'] for d in dataconfset]
            if len(set(types)) > 1:
                raise Exception('all data types in a set must be the same')
            self.readers.append(tfreaders.factory(types[0])([d['dir'] for d in dataconfset]))
        hist = self.readers[0].metadata['sequence_length_histogram']
        self.max_length = hist.size
        self.boundaries, self.batch_sizes, self.num_steps = batch_plan(hist, batch_size, numbuckets, variable_batch_size)
        self.allow_smaller_final_batch = allow_smaller_final_batch
        self.shuffle_seed = shuffle_seed
        self.input_dims = {n: self.readers[i].metadata['dim'] for i, n in enumerate(self.input_names)
                           if 'dim' in self.readers[i].metadata}
        self._epoch = 0
        # synchronous data parallelism: every rank walks the SAME global batches (same seed, same order)
        # and keeps utterances rank::world of each; equal shard sizes are what makes the mean of the ranks' batch means
        # the global batch mean, so every batch size must divide
        # batches read, parsed and padded ahead of the consumer by ONE background thread (the reference's queue runners,
        # processing/input_pipeline.py:133-168; one thread keeps the order deterministic); 0 = read in the caller
        self.prefetch = int(os.environ.get('NABU_PREFETCH', '0')) if prefetch is None else int(prefetch)
        self.rank, self.world = int(rank), int(world)
        if self.world > 1:
            # The reference's per-bucket sizes (16, 14, 13, 11, ... with variable_batch_size) do not divide the number of
            # ranks in general: round every bucket's size DOWN to a multiple of `world` (at least one utterance per
            # rank) and recount the steps from the rounded sizes, so that stock recipes train under torchrun unchanged.
            rounded = [max(b // self.world, 1) * self.world for b in self.batch_sizes]
            if rounded != self.batch_sizes:
                edges = [0] + list(self.boundaries) + [hist.size]
                numutt = [hist[edges[i]:edges[i + 1]].sum() for i in range(len(rounded))]
                self.batch_sizes = rounded
                self.num_steps = int(sum(int(n) // b for n, b in zip(numutt, rounded)))

    def __len__(self):
        return self.num_steps

    def _bucket(self, length):
        # bucket_by_sequence_length: bucket i holds lengths in [boundaries[i-1], boundaries[i])
        return int(np.searchsorted(np.asarray(self.boundaries), length, side='right')) if self.boundaries else 0

    def _host_batch(self, items, pin=False):
        """padded host tensors of one batch; `pin`: page-locked when they are going to a GPU, so that the copy can be
        asynchronous (the prefetch thread's batches)"""
        torch = self._torch
        if self.world > 1:
            items = items[self.rank::self.world]
        streams = list(zip(*items))                     # per stream: list of (array, length)
        pin = pin and torch.device(self.device).type == 'cuda' and torch.cuda.is_available()
        tensors, lengths = [], []
        for st in streams:
            t = torch.from_numpy(_pad([a for a, _ in st]))
            n = torch.tensor([l for _, l in st], dtype=torch.int32)
            tensors.append(t.pin_memory() if pin else t)
            lengths.append(n.pin_memory() if pin else n)
        return tensors, lengths

    def _emit(self, items):
        return self._to_device(*self._host_batch(items))

    def _to_device(self, tensors, lengths):
        tensors = [t.to(self.device, non_blocking=t.is_pinned()) for t in tensors]
        lengths = [n.to(self.device, non_blocking=n.is_pinned()) for n in lengths]
        ni = len(self.input_names)
        inputs = {n: tensors[i] for i, n in enumerate(self.input_names)}
        ilen = {n: lengths[i] for i, n in enumerate(self.input_names)}
        targets = {n: tensors[ni + i] for i, n in enumerate(self.target_names)}
        tlen = {n: lengths[ni + i] for i, n in enumerate(self.target_names)}
        return inputs, ilen, targets, tlen

    def __iter__(self):
        if self.prefetch <= 0:
            for items in self._batches():
                yield self._emit(items)
            return
        import queue
        import threading
        q = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def produce():
            try:
                for items in self._batches():
                    batch = self._host_batch(items, pin=True)
                    while not stop.is_set():
                        try:
                            q.put(batch, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                    if stop.is_set():
                        return
                q.put(None)
            except BaseException as e:                   # surfaces in the consumer, not in a dead thread
                q.put(e)

        worker = threading.Thread(target=produce, daemon=True)
        worker.start()
        try:
            while True:
                batch = q.get()
                if batch is None:
                    return
                if isinstance(batch, BaseException):
                    raise batch
                yield self._to_device(*batch)
        finally:
            stop.set()

    def _batches(self):
        """the examples of one pass over the data, grouped into batches (lists of per-stream (array, length))"""
        order = list(range(len(self.elements)))
        if self.shuffle_seed is not None:
            np.random.default_rng(self.shuffle_seed + self._epoch).shuffle(order)
        self._epoch += 1
        buckets = [[] for _ in self.batch_sizes]
        for idx in order:
            files = self.elements[idx].split('\t')
            item = [reader(f) for reader, f in zip(self.readers, files)]
            b = self._bucket(item[0][1])
            buckets[b].append(item)
            if len(buckets[b]) == self.batch_sizes[b]:
                yield buckets[b]
                buckets[b] = []
        if self.allow_smaller_final_batch:
            for items in buckets:
                # data parallel: a tail smaller than a multiple of the ranks is cut to that multiple (every rank needs
                # the same number of utterances for the mean of the ranks' means to be the batch mean); fewer
                # utterances than ranks are dropped
                if self.world > 1:
                    items = items[:len(items) // self.world * self.world]
                if items:
                    yield items


def dataconfs_for(conf, dataconf, names):
    """The database sections a trainer / evaluator / recognizer cfg names for each stream: `conf[name]` is a space
    separated list of sections of the database cfg (trainers/trainer.py:303-326, evaluators/evaluator.py:38-55,
    recognizer.py:42-48)."""
    out = []
    for name in names:
        out.append([dict(dataconf.items(section)) for section in conf[name].split(' ')])
    return out


def source_from_conf(conf, dataconf, input_names, target_names, device='cpu', **kw):
    """BatchSource over the data sections `conf` names (trainer: shuffled, bucketed, variable batch size as its cfg
    says; evaluator / recognizer: one bucket, their own batch_size)."""
    confs = dataconfs_for(conf, dataconf, list(input_names) + list(target_names))
    return BatchSource(confs, input_names, target_names, batch_size=int(conf['batch_size']), device=device, **kw)
