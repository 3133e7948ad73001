"""Host -> device batch pipeline for the training / evaluation loops (last item).

The reference moves every batch with `batch = [elt.to(device) for elt in batch]` on the compute stream from PAGEABLE memory
(reference models/model.py:229, 424; its DataLoaders are built with `pin_memory=False`, reference functions.py:172, 197): a blocking
staged copy in front of every step.  `DeviceBatchPrefetcher` wraps the same iterable (the reference's DataLoader or any iterator of
tensor tuples) and yields the same batches already on the device:

    for step, batch in enumerate(DeviceBatchPrefetcher(dataset_train, device)):     # instead of: batch = [elt.to(device) ...]
        pred = model.forward(batch) ...

  * each host tensor is pinned (a no-op when the DataLoader pins already) and copied on a dedicated copy stream while the previous
    step computes (`depth` batches in flight, default 2);
  * the consumer's stream waits on the copy's event only (no host synchronisation), and every yielded tensor is registered with that
    stream (`record_stream`) so the caching allocator cannot hand its memory to a later copy while the step still reads it;
  * non-tensor entries pass through unchanged, order and values are the DataLoader's.
There is no CPU mode: the pipeline exists to feed the CUDA path and raises for any other device."""
import collections

import torch


class DeviceBatchPrefetcher:
    def __init__(self, batches, device, depth: int = 2):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceBatchPrefetcher feeds the CUDA path only (the reference's --cpu route needs no transfer)")
        if depth < 1:
            raise ValueError("depth must be at least 1")
        self.batches, self.depth = batches, int(depth)
        self._stream = None

    def __len__(self):
        return len(self.batches)

    def _stage(self, batch):
        """Issue the copies of one batch on the copy stream -> (device batch, event)."""
        out = []
        with torch.cuda.stream(self._stream):
            for elt in batch:
                if torch.is_tensor(elt):
                    src = elt if (elt.is_cuda or elt.is_pinned()) else elt.pin_memory()
                    out.append(src.to(self.device, non_blocking=True))
                else:
                    out.append(elt)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return (tuple(out) if isinstance(batch, tuple) else out), ev

    def __iter__(self):
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=self.device)
        queue = collections.deque()
        it = iter(self.batches)
        done = False
        while True:
            while not done and len(queue) < self.depth:
                try:
                    queue.append(self._stage(next(it)))
                except StopIteration:
                    done = True
            if not queue:
                return
            batch, ev = queue.popleft()
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            for elt in batch:
                if torch.is_tensor(elt) and elt.is_cuda:
                    elt.record_stream(cur)
            yield batch
