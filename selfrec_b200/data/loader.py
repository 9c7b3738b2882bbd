"""Minimal mirror of data/loader.py (FileIO): `user item weight` triples per line."""
import os


class FileIO(object):
    @staticmethod
    def write_file(dir, file, content, op="w"):
        os.makedirs(dir, exist_ok=True)
        with open(os.path.join(dir, file), op) as f:
            f.writelines(content)

    @staticmethod
    def delete_file(file_path):
        if os.path.exists(file_path):
            os.remove(file_path)

    @staticmethod
    def load_data_set(file, rec_type="graph"):
        if rec_type != "graph":
            raise NotImplementedError("selfrec_b200 covers the graph models only")
        data = []
        with open(file) as f:
            for line in f:
                parts = line.strip().split(" ")
                data.append([parts[0], parts[1], float(parts[2])])
        return data
