"""TEST INFRASTRUCTURE — CPU oracle of the region-diffusion hot path (restatement of the reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package. The product package (rich-text-to-image_b200/, alias rtti_b200) never does.
"""
