"""Writes tests/golden/muller_2022.json: the muller_2022 example network (a DATA file of the
reference, test/example_networks/muller_2022.phy) together with the structure sizes that the
reference's jldoctests record for it (docs/src/man/clustergraphs.md:40-41,52-56,101-117,131-144,198-211).
Needs /root/reference (build container only); the fixture travels with the repository."""
import json
import os

REF = "/root/reference"
net = open(os.path.join(REF, "test", "example_networks", "muller_2022.phy")).read().strip()
G = {
    "_source": "JuliaPhylo/PhyloGaussianBeliefProp.jl v0.0.1 test/example_networks/muller_2022.phy + docs/src/man/clustergraphs.md",
    "newick": net,
    # docs/src/man/clustergraphs.md:40-41
    "nnodes": 801, "ntips": 40, "nhybrids": 361, "nedges": 1161,
    # :52-56 clique tree; :101-105 Bethe (cluster size <= 3, mean 1.743738); :198-211 LTRIP on the node families
    "cliquetree": {"nclusters": 664, "nsepsets": 663},
    "bethe": {"nclusters": 1557, "nsepsets": 1914, "max_cluster_size": 3, "mean_cluster_size": 1.743738},
}
json.dump(G, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "muller_2022.json"), "w"), indent=1)
print("ok", len(net))
