import json

import numpy as np


class JsonCustomEncoder(json.JSONEncoder):
    def default(self, obj):
        if hasattr(obj, 'to_string'):
            return obj.to_string()
        if isinstance(obj, (np.number, np.ndarray)):
            return obj.tolist()
        return json.JSONEncoder.default(self, obj)
