"""CLI entry with the reference's contract:  python main_from_args.py <args.json> <DATASET> <split/>
(run/main_from_args.py:79-98 of nju-websoft/OpenEA).  Model names resolve exactly as there."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from openea.modules.args.args_hander import check_args, load_args  # noqa: E402,F401
from openea.modules.load.kgs import read_kgs_from_folder  # noqa: E402
from openea.models.trans import TransD, TransE, TransH, TransR  # noqa: E402
from openea.models.semantic import DistMult, HolE, SimplE, RotatE  # noqa: E402
from openea.models.neural import ConvE, ProjE  # noqa: E402
from openea.approaches import (AlignE, BootEA, JAPE, Attr2Vec, MTransE, IPTransE, GCN_Align, AttrE, IMUSE, SEA,  # noqa: E402
                               MultiKE, RSN4EA, GMNN, KDCoE, RDGCN, BootEA_RotatE, BootEA_TransH, AliNet)
from openea.models.basic_model import BasicModel  # noqa: E402

_MODELS = dict(BasicModel=BasicModel, TransE=TransE, TransD=TransD, TransH=TransH, TransR=TransR, DistMult=DistMult,
               HolE=HolE, SimplE=SimplE, RotatE=RotatE, ProjE=ProjE, ConvE=ConvE, MTransE=MTransE, IPTransE=IPTransE,
               Attr2Vec=Attr2Vec, JAPE=JAPE, AlignE=AlignE, BootEA=BootEA, GCN_Align=GCN_Align, GMNN=GMNN, KDCoE=KDCoE,
               AttrE=AttrE, IMUSE=IMUSE, SEA=SEA, MultiKE=MultiKE, RSN4EA=RSN4EA, RDGCN=RDGCN,
               BootEA_RotatE=BootEA_RotatE, BootEA_TransH=BootEA_TransH, AliNet=AliNet)


class ModelFamily(object):
    pass


for _name, _cls in _MODELS.items():
    setattr(ModelFamily, _name, _cls)


def get_model(model_name):
    return getattr(ModelFamily, model_name)


def init_process_group_from_env():
    """Under `torchrun --nproc-per-node G main_from_args.py …` (one process per GPU): bind this rank to its GPU and join
    the NCCL group; the GNN approaches then row-shard their graphs (openea_b200/parallel_gnn.py).  A plain
    `python main_from_args.py …` run is untouched."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))


if __name__ == '__main__':
    t = time.time()
    init_process_group_from_env()
    args = load_args(sys.argv[1])
    args.training_data = args.training_data + sys.argv[2] + '/'
    args.dataset_division = sys.argv[3]
    print(args.embedding_module)
    print(args)
    remove_unlinked = args.embedding_module == "RSN4EA"
    kgs = read_kgs_from_folder(args.training_data, args.dataset_division, args.alignment_module, args.ordered,
                               remove_unlinked=remove_unlinked)
    model = get_model(args.embedding_module)()
    model.set_args(args)
    model.set_kgs(kgs)
    model.init()
    model.run()
    model.test()
    model.save()
    print("Total run time = {:.3f} s.".format(time.time() - t))
