from .dataload import Data, SyntheticData
from .dataset import SEQTrainDataset, SeqEvalDataset, seq_eval_collate
from .utils import bulid_dataloader, load_data

__all__ = ["load_data", "bulid_dataloader", "Data", "SyntheticData", "SEQTrainDataset", "SeqEvalDataset", "seq_eval_collate"]
