"""Pipeline-parallel engine: the surface train.py / utils/dataset.py / utils/saver.py use from DeepSpeed
(SURVEY.md section 8b, B-py.2), re-implemented over the C++ 1F1B planner and torch.distributed / CUDA-IPC links."""
from .module import LayerSpec, ManualPipelineModule, PipelineModule  # noqa: F401
from .engine import PipelineEngine, initialize  # noqa: F401
from .schedule import (BackwardPass, ForwardPass, InferenceSchedule, LoadMicroBatch, OptimizerStep, RecvActivation,  # noqa: F401
                       RecvGrad, ReduceGrads, ReduceTiedGrads, SendActivation, SendGrad, TrainSchedule, ZeroBubbleSchedule,
                       BackwardInput, BackwardWeight)
