"""Shape / batch constants of the hot path (mirror of reference shared/param.py:9-16) and the training hyper-parameters
the training step reads (shared/param.py:15-25)."""
flankingBaseNum = 16          # shared/param.py:9
matrixRow = 8                 # shared/param.py:10
matrixNum = 4                 # shared/param.py:11
predictBatchSize = 1000       # shared/param.py:16
expandReferenceRegion = 1000000      # shared/param.py:5 (CreateTensor.py:131)
SAMTOOLS_VIEW_FILTER_FLAG = 2316     # shared/param.py:6 (CreateTensor.py:174)
trainBatchSize = 10000        # shared/param.py:15
initialLearningRate = 1e-3    # shared/param.py:17
learningRateDecay = 0.1        # shared/param.py:18 (applied by the training loop, clair/train.py:201-215)
l2RegularizationLambda = 0.005    # shared/param.py:23
NUM_THREADS = 12              # shared/param.py:3 (kept for callers that mutate it: call_var.py:182-189)

no_of_positions = 2 * flankingBaseNum + 1
input_tensor_size = no_of_positions * matrixRow * matrixNum   # clair/utils.py:68-69


def get_model_parameters():   # shared/param.py:51-56
    return dict(flankingBaseNum=flankingBaseNum, matrixNum=matrixNum, expandReferenceRegion=expandReferenceRegion)
