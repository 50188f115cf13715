"""The batch loop that drives predict(): same contract as ``clair.call_var.call_variants``
(reference clair/call_var.py:1312-1367).

The reference runs three stages per iteration and joins them at a barrier:
  output(batch k-1)  ||  predict(batch k)  ||  load(batch k+1)
exactly one predict is in flight, batches are handled in strict input order, and the output stage
of an iteration is handed the ``m.prediction`` object that existed when the iteration was set up
(call_var.py:1334-1338) - i.e. the result of the previous iteration's predict.  ``run_batches(..., in_flight=1)``
is that loop, stage for stage.

One 1000-site predict (shared/param.py:16) fills 8 of a B200's 74 CTA pairs, so the default loop keeps the same three
stages and the same order but lets them run ahead of each other: a loader thread, the submitting thread
(``m.predict_async``: up to ``in_flight`` batches queued, packed into full device chunks by the library) and an output
thread that receives every batch with exactly its own prediction, in input order.
The VCF decision logic (batch_output, call_var.py:1199-1236) is out of scope and is injected.
"""
import logging
import queue
from threading import Semaphore, Thread
from time import time

from . import param, utils

DEFAULT_IN_FLIGHT = 64      # predict-batches between the loader and the output stage (>= 3 device chunks of 18,944 sites)


def run_batches(m, tensor_generator, output_stage, *output_args, with_decision=False, in_flight=None, release=None):
    """Drive ``m.predict`` over every (X, infos) batch of ``tensor_generator``.

    with_decision=True drives ``m.predict_and_decide`` instead (reference bases from the info triples,
    clair_b200.decision.ref_base_codes) and leaves the decision records of the batch in ``m.decision`` next to
    ``m.prediction``; the output stage is then called as ``output_stage(batch, prediction, decision, *output_args)``
    with the records of exactly that batch (the same one-iteration hand-over as the prediction).

    in_flight: predict-batches that may be queued on the device at once; None = DEFAULT_IN_FLIGHT when the model has
    ``predict_async``, else 1.  in_flight=1 is the reference's lock-step loop.  release(X), if given, is called once
    the output stage of a batch has returned (hands a pinned staging buffer back to its pool)."""
    if in_flight is None:
        in_flight = DEFAULT_IN_FLIGHT if hasattr(m, "predict_async") else 1
    if in_flight > 1:
        if not hasattr(m, "predict_async"):
            raise ValueError("in_flight > 1 needs a model with predict_async")
        return _run_pipelined(m, tensor_generator, output_stage, output_args, with_decision, in_flight, release)
    to_predict = None      # batch loaded in the previous iteration
    to_output = None       # batch predicted in the previous iteration
    source_open = True
    if with_decision:
        from . import decision as _decision

        def predict_stage(batch):
            _, dec = m.predict_and_decide(batch[0], _decision.ref_base_codes(batch[1]))
            m.decision = dec

    while True:
        stages = []
        if to_output is not None:
            handed = (to_output, m.prediction, m.decision) if with_decision else (to_output, m.prediction)
            stages.append(Thread(target=output_stage, args=handed + output_args))
        predicted = None
        if to_predict is not None:
            if with_decision:
                stages.append(Thread(target=predict_stage, args=(to_predict,)))
            else:
                stages.append(Thread(target=m.predict, kwargs={"batchX": to_predict[0]}))
            predicted = to_predict
        fetched = []
        if source_open:
            stages.append(Thread(target=lambda: fetched.extend(_take_one(tensor_generator))))
        if not stages:
            return
        for s in stages:
            s.start()
        for s in stages:
            s.join()
        if release is not None and to_output is not None:
            release(to_output[0])
        to_output = predicted
        to_predict = fetched[0] if fetched else None
        if source_open and not fetched:
            source_open = False


_END = object()


def _run_pipelined(m, tensor_generator, output_stage, output_args, with_decision, in_flight, release):
    """load -> predict_async -> output as three free-running stages joined by bounded queues (order preserved)."""
    if with_decision:
        from . import decision as _decision
    loaded = queue.Queue(maxsize=max(2, in_flight // 2))
    submitted = queue.Queue()
    slots = Semaphore(in_flight)                               # tickets between predict_async and result()
    failure = []

    def load_stage():
        try:
            for batch in tensor_generator:
                if failure:
                    break
                loaded.put(batch)
        except BaseException as exc:
            failure.append(exc)
        finally:
            loaded.put(_END)

    def out_stage():
        while True:
            item = submitted.get()
            if item is _END:
                return
            batch, ticket = item
            try:
                try:
                    result = ticket.result()                   # also on the way out: every ticket is waited for
                finally:
                    slots.release()
                if failure:
                    continue
                if with_decision:
                    prediction, dec = result
                    m.decision = dec
                    output_stage(batch, prediction, dec, *output_args)
                else:
                    output_stage(batch, result, *output_args)
                if release is not None:
                    release(batch[0])
            except BaseException as exc:
                failure.append(exc)

    loader, writer = Thread(target=load_stage), Thread(target=out_stage)
    loader.start()
    writer.start()
    try:
        while True:
            batch = loaded.get()
            if batch is _END:
                break
            if failure:
                continue                                       # keep draining the loader so it can finish
            slots.acquire()
            try:
                if with_decision:
                    ticket = m.predict_async(batch[0], _decision.ref_base_codes(batch[1]))
                else:
                    ticket = m.predict_async(batch[0])
            except BaseException as exc:
                slots.release()
                failure.append(exc)
                continue
            submitted.put((batch, ticket))
    finally:
        submitted.put(_END)
        loader.join()
        writer.join()
    if failure:
        raise failure[0]


def _take_one(gen):
    try:
        return [next(gen)]
    except StopIteration:
        return []


def call_variants_from_alignments(block, m, output_config, output_utilities, output_stage):
    """call_variants fed by CreateTensor on the device (clair_b200.create_tensor.create_tensors(..., subtract=True)) instead of
    its text rows: same batch loop, same (X, infos) hand-over, the tensors never leave the GPU - the pipe between the two
    reference processes (clair/callVarBam.py:191-200) is gone.  An output stage that reads the tensors themselves, as the
    reference's batch_output does for read depth and supporting reads (clair/call_var.py:1021-1151), iterates the batch's
    DeviceTensors: create the block with fetch=True then (one device->host copy; the forward still reads the resident block)."""
    from . import create_tensor
    output_utilities.output_header()
    logging.info("Calling variants ...")
    started = time()
    run_batches(m, create_tensor.device_tensor_generator_from(block, param.predictBatchSize), output_stage, output_config,
                output_utilities)
    logging.info("Total time elapsed: %.2f s" % (time() - started))
    output_utilities.close_opened_files()


def call_variants(args, m, output_config, output_utilities, output_stage, in_flight=None):
    """Reference signature plus the output stage to run (the reference picks batch_output or
    batch_output_for_ensemble itself, call_var.py:1320).  Batches are decoded straight into a ring of pinned staging
    buffers (the host->device copies are then asynchronous) and, with a model that has predict_async, kept `in_flight`
    deep on the device."""
    from .model import PinnedPool
    output_utilities.output_header()
    if in_flight is None:
        in_flight = DEFAULT_IN_FLIGHT if hasattr(m, "predict_async") else 1
    pool = PinnedPool(2 * in_flight + 4, (param.predictBatchSize, utils.input_tensor_size), "float32")
    try:
        tensor_generator = utils.tensor_generator_from(args.tensor_fn, param.predictBatchSize, alloc=pool.take)
        logging.info("Calling variants ...")
        started = time()
        run_batches(m, tensor_generator, output_stage, output_config, output_utilities, in_flight=in_flight, release=pool.give)
        logging.info("Total time elapsed: %.2f s" % (time() - started))
    finally:
        pool.close()
    output_utilities.close_opened_files()
