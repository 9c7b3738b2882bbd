"""Ranking metrics with the output format of util/evaluation.py:135-162.

ranking_evaluation(origin, res, N) keeps the reference's string protocol (fast_evaluation
re-parses it, graph_recommender.py:84-86): for each n a 'Top n' line followed by
'Hit Ratio:', 'Precision:', 'Recall:', 'NDCG:' lines, values rounded to 5 decimals.
"""
import math


def _hits(origin, predicted):
    return {u: len(set(origin[u]).intersection(it[0] for it in predicted[u])) for u in origin}


def ranking_evaluation(origin, res, N):
    measure = []
    for n in N:
        predicted = {u: res[u][:n] for u in res}
        if len(origin) != len(predicted):
            print("The Lengths of test set and predicted set do not match!")
            exit(-1)
        hits = _hits(origin, predicted)
        total = sum(len(origin[u]) for u in origin)
        hit_sum = sum(hits.values())
        hr = round(hit_sum / total, 5)
        prec = round(hit_sum / (len(hits) * n), 5)
        rec_list = [hits[u] / len(origin[u]) for u in hits]
        recall = round(sum(rec_list) / len(rec_list), 5)
        ndcg_sum = 0
        for u in predicted:
            dcg = sum(1.0 / math.log(r + 2, 2) for r, it in enumerate(predicted[u]) if it[0] in origin[u])
            idcg = sum(1.0 / math.log(r + 2, 2) for r in range(min(len(origin[u]), n)))
            ndcg_sum += dcg / idcg
        ndcg = round(ndcg_sum / len(predicted), 5)
        measure.append("Top " + str(n) + "\n")
        measure += ["Hit Ratio:" + str(hr) + "\n", "Precision:" + str(prec) + "\n", "Recall:" + str(recall) + "\n",
                    "NDCG:" + str(ndcg) + "\n"]
    return measure
