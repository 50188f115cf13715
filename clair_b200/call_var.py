"""The batch loop that drives predict(): same contract as ``clair.call_var.call_variants``
(reference clair/call_var.py:1312-1367).

Per iteration three stages run concurrently and then meet at a barrier:
  output(batch k-1)  ||  predict(batch k)  ||  load(batch k+1)
exactly one predict is in flight, batches are handled in strict input order, and the output stage
of an iteration is handed the ``m.prediction`` object that existed when the iteration was set up
(call_var.py:1334-1338) - i.e. the result of the previous iteration's predict.
The VCF decision logic (batch_output, call_var.py:1199-1236) is out of scope and is injected.
"""
import logging
from threading import Thread
from time import time

from . import param, utils


def run_batches(m, tensor_generator, output_stage, *output_args, with_decision=False):
    """Drive ``m.predict`` over every (X, infos) batch of ``tensor_generator``.

    with_decision=True drives ``m.predict_and_decide`` instead (reference bases from the info triples,
    clair_b200.decision.ref_base_codes) and leaves the decision records of the batch in ``m.decision`` next to
    ``m.prediction``; the output stage is then called as ``output_stage(batch, prediction, decision, *output_args)``
    with the records of exactly that batch (the same one-iteration hand-over as the prediction)."""
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
        to_output = predicted
        to_predict = fetched[0] if fetched else None
        if source_open and not fetched:
            source_open = False


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


def call_variants(args, m, output_config, output_utilities, output_stage):
    """Reference signature plus the output stage to run (the reference picks batch_output or
    batch_output_for_ensemble itself, call_var.py:1320)."""
    output_utilities.output_header()
    tensor_generator = utils.tensor_generator_from(args.tensor_fn, param.predictBatchSize)
    logging.info("Calling variants ...")
    started = time()
    run_batches(m, tensor_generator, output_stage, output_config, output_utilities)
    logging.info("Total time elapsed: %.2f s" % (time() - started))
    output_utilities.close_opened_files()
