from openea_b200.models import trans
from openea_b200.models import semantic
from openea_b200.models import neural
from openea_b200.models import attr
