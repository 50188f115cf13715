"""The second caller of Clair.predict: evaluate_model (reference clair/evaluate.py:37-148).

Same walk over the frames of a bin (clair_b200.bins), same `m.predict(x_batch)` per 1000 sites, same report on stdout:
top-1 / top-2 counts of the 21-genotype head, the four confusion matrices and their per-class f-measures.  The per-site
Python loops of the reference (one arg-max and one matrix increment per site and task) are numpy reductions over the batch
here.  Pinned: tests/golden/evaluate_report.txt is what the reference's own evaluate_model printed for a seeded bin and a
deterministic stand-in model (oracle/gen_golden_evaluate.py); this function prints the same text.
"""
import logging
from time import time

import numpy as np

from . import bins, param

# (labels, first column, one past the last column) of the four tasks in a label row (clair/task/main.py:10-29)
GT21, GENOTYPE, VARIANT_LENGTH_1, VARIANT_LENGTH_2 = (21, 0, 21), (3, 21, 24), (33, 24, 57), (33, 57, 90)


def f1_score(confusion_matrix):
    """clair/evaluate.py:18-31, all classes at once"""
    epsilon = 1e-15
    tp = np.diagonal(confusion_matrix) + 0.0
    precision = tp / (confusion_matrix.sum(axis=0) + epsilon)
    recall = tp / (confusion_matrix.sum(axis=1) + epsilon)
    return (2.0 * precision * recall) / (precision + recall + epsilon)


def _count(matrix, true_index, predicted_index):
    np.add.at(matrix, (true_index, predicted_index), 1)


def evaluate_model(m, dataset_info):
    batch_size = param.predictBatchSize
    no_of_training_examples = dataset_info.no_of_training_examples_from_train_binary or \
        int(dataset_info.dataset_size * bins.trainingDatasetPercentage)
    no_of_blosc_blocks = bins.no_of_blosc_blocks_from(dataset_info, no_of_training_examples, bins.bloscBlockSize)
    logging.info("[INFO] Testing on the training and validation dataset ...")
    started = time()
    matrices = [np.zeros((task[0], task[0]), dtype=np.int64) for task in (GT21, GENOTYPE, VARIANT_LENGTH_1, VARIANT_LENGTH_2)]
    all_gt21_count = top_1_count = top_2_count = 0
    blosc_index, first_row = 0, 0
    while True:
        x_batch, next_first_row, next_blosc_index = bins.decompress_array(
            dataset_info.x_array_compressed, blosc_index, first_row, batch_size, no_of_blosc_blocks)
        y_batch, _, _ = bins.decompress_array(dataset_info.y_array_compressed, blosc_index, first_row, batch_size, no_of_blosc_blocks)
        gt21, genotype, length_1, length_2 = m.predict(x_batch)
        blosc_index, first_row = next_blosc_index, next_first_row

        def label(task):
            return np.argmax(y_batch[:, task[1]:task[2]], axis=1)

        true_gt21 = label(GT21)
        _count(matrices[0], true_gt21, np.argmax(gt21, axis=1))
        # top-2: the two largest probabilities in the order argsort()[::-1] gives them (clair/evaluate.py:97-102)
        order = np.argsort(gt21, axis=1)[:, ::-1]
        hit_1 = order[:, 0] == true_gt21
        all_gt21_count += len(true_gt21)
        top_1_count += int(hit_1.sum())
        top_2_count += int((hit_1 | (order[:, 1] == true_gt21)).sum())
        _count(matrices[1], label(GENOTYPE), np.argmax(genotype, axis=1))
        # the two lengths are compared as an unordered pair: (smaller, larger) on both sides (:121-127)
        true_pair = np.sort(np.stack([label(VARIANT_LENGTH_1), label(VARIANT_LENGTH_2)], axis=1), axis=1)
        predicted_pair = np.sort(np.stack([np.argmax(length_1, axis=1), np.argmax(length_2, axis=1)], axis=1), axis=1)
        _count(matrices[2], true_pair[:, 0], predicted_pair[:, 0])
        _count(matrices[3], true_pair[:, 1], predicted_pair[:, 1])
        if not (next_first_row >= 0 and next_blosc_index >= 0):
            break
    logging.info("[INFO] Prediciton time elapsed: %.2f s" % (time() - started))

    def print_matrix(matrix):
        for row in matrix:
            print("\t".join(str(v) for v in row))
        print("[INFO] f-measure: ", f1_score(matrix))

    print("[INFO] Evaluation on gt21:")
    print("[INFO] all/top1/top2/top1p/top2p: %d/%d/%d/%.2f/%.2f" % (
        all_gt21_count, top_1_count, top_2_count,
        float(top_1_count) / all_gt21_count * 100, float(top_2_count) / all_gt21_count * 100))
    print_matrix(matrices[0])
    print("\n[INFO] Evaluation on Genotype:")
    print_matrix(matrices[1])
    print("\n[INFO] evaluation on indel length 1:")
    print_matrix(matrices[2])
    print("\n[INFO] evaluation on indel length 2:")
    print_matrix(matrices[3])
    return matrices
